#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for lay in thin3 fat3; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "collect_layouts and $lay" > $O/r02c_layout_$lay.txt 2>&1
  echo "layout $lay: rc=$? $(tail -1 $O/r02c_layout_$lay.txt)"
done
rm -f $O/r02c_timing.txt
for lay in sets thin3 thin2 fat3; do
  echo "== FWAV_UMMA_COLLECT=$lay (config 2)" >> $O/r02c_timing.txt
  FWAV_UMMA_COLLECT=$lay timeout 200 python scripts/time_topk.py 1.0 umma 3 2>/dev/null | cut -c1-400 >> $O/r02c_timing.txt
done
cat $O/r02c_timing.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_batch or decode" > $O/r02c_mb.txt 2>&1
echo "multi-batch + decode: rc=$? $(tail -1 $O/r02c_mb.txt)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r02c_bench.json 2> $O/r02c_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'], d['kernels'], d['decode'])
PY
