#!/bin/bash
# top_k = 64 failure routes: parity tests, then the config-4 shape at 1/10 length with the per-batch log
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "search_paths or odd_shapes" 2>&1 | tail -15 > gpurun_out/retry_tests.txt
cat gpurun_out/retry_tests.txt
FWAV_UMMA_VERBOSE=1 timeout 300 python bench.py --workload c4 --scale ${1:-0.1} --steps 1 --warmup 1 --no-decode --no-cpu > gpurun_out/bench_c4_s.json 2> gpurun_out/bench_c4_s.err
echo "exit $?"; grep "fwav\]" gpurun_out/bench_c4_s.err | tail -12 | cut -c1-260
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c4_s.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline'].get('search_phases_ms'))
PY
