#!/bin/bash
# per-kernel durations of the search at config-2 scale with parts of the epilogue switched off
mkdir -p gpurun_out; rm -f gpurun_out/dbg6.txt
for d in ${DBGS:-0 4 10}; do
  echo "== FWAV_UMMA_DEBUG=$d" >> gpurun_out/dbg6.txt
  FWAV_UMMA_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dbg6_$d.csv python scripts/time_topk.py ${1:-1.0} umma 0 > gpurun_out/dbg6_ncu.log 2>&1
  python - $d >> gpurun_out/dbg6.txt <<'PY'
import csv,collections,sys
rows=[r for r in csv.reader(open(f'gpurun_out/dbg6_{sys.argv[1]}.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-40:]; v=float(r[-1].replace(',',''))
    agg.setdefault(name,[]).append(v)
for n,v in agg.items():
    if 'scan' in n or 'final' in n: print(f"{n:42s} n={len(v):2d} each={sum(v)/len(v)/1e6:9.3f} ms")
PY
done
cat gpurun_out/dbg6.txt
