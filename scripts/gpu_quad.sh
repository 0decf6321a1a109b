#!/bin/bash
# Round-2 opener for the two experimental paths (four-buffer collect pass, compact split): parity first, then the
# config-2 search with and without it, then the config-4 shape at 1/10 length (full split).
mkdir -p gpurun_out; rm -f gpurun_out/quad.txt
FWAV_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental" 2>&1 | tail -5 > gpurun_out/quad_tests.txt
cat gpurun_out/quad_tests.txt
grep -q "passed" gpurun_out/quad_tests.txt || exit 1
# the N=8 fixture that has not been on a device yet (tile 2048)
FWAV_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "music_t2048" 2>&1 | tail -4
for quad in 0 1; do
  echo "== FWAV_UMMA_QUAD=$quad (config 2)" >> gpurun_out/quad.txt
  FWAV_UMMA_QUAD=$quad FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > gpurun_out/quad_run.out 2> gpurun_out/quad_run.err
  grep "fwav\]" gpurun_out/quad_run.err | tail -3 | cut -c1-200 >> gpurun_out/quad.txt
  cut -c1-260 gpurun_out/quad_run.out >> gpurun_out/quad.txt
  echo "== FWAV_UMMA_QUAD=$quad FWAV_UMMA_MODE=precise (config 2, full split)" >> gpurun_out/quad.txt
  FWAV_UMMA_QUAD=$quad FWAV_UMMA_MODE=precise timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-260 >> gpurun_out/quad.txt
done
for compact in 0 1; do
  echo "== FWAV_UMMA_COMPACT=$compact (config-4 shape at 1/10 length)" >> gpurun_out/quad.txt
  FWAV_UMMA_COMPACT=$compact FWAV_UMMA_VERBOSE=1 timeout 200 python bench.py --workload c4 --scale 0.1 --steps 1 --warmup 1 --no-decode --no-cpu > gpurun_out/quad_c4_$compact.json 2> gpurun_out/quad_c4_$compact.err
  grep "live embedding" gpurun_out/quad_c4_$compact.err | tail -1 | cut -c1-200 >> gpurun_out/quad.txt
  python - $compact >> gpurun_out/quad.txt <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/quad_c4_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline'].get('search_phases_ms'))
PY
done
cat gpurun_out/quad.txt
