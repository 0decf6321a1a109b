#!/bin/bash
# per-kernel durations of one search call at config-2 scale (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
timeout 200 python scripts/time_topk.py ${1:-1.0} umma 1 > gpurun_out/topk_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/topk_launches.csv python scripts/time_topk.py ${1:-1.0} umma 1 > gpurun_out/topk_ncu.log 2>&1
tail -1 gpurun_out/topk_plain.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/topk_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-60:]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    agg.setdefault((name,unit),[]).append(v)
for (n,u),v in agg.items(): print(f"{n:62s} n={len(v):3d} total={sum(v):14.1f} {u} each={sum(v)/len(v):12.1f}")
PY
