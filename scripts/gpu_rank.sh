#!/bin/bash
# threshold-rank sweep: config-2 search alone, then the config-4 shape at 1/10 length
mkdir -p gpurun_out; rm -f gpurun_out/rank.txt
for r in 16 12 10; do
  FWAV_UMMA_RANK=$r FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 >> gpurun_out/rank.txt 2> gpurun_out/rank_$r.err
  grep "fwav\]" gpurun_out/rank_$r.err | tail -2 | cut -c1-200 >> gpurun_out/rank.txt
done
for r in 24 18; do
  FWAV_UMMA_RANK=$r FWAV_UMMA_VERBOSE=1 timeout 200 python bench.py --workload c4 --scale 0.1 --steps 1 --warmup 1 --no-decode --no-cpu > gpurun_out/rank_c4_$r.json 2> gpurun_out/rank_c4_$r.err
  grep "fwav\]" gpurun_out/rank_c4_$r.err | tail -3 | cut -c1-200 >> gpurun_out/rank.txt
  python - $r >> gpurun_out/rank.txt <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/rank_c4_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('c4@0.1 rank', sys.argv[1], d['ms_per_step'], d['roofline'].get('search_phases_ms'))
PY
done
cat gpurun_out/rank.txt
