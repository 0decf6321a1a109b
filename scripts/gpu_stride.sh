#!/bin/bash
# sample stride x threshold rank sweep of the config-2 search (after the search parity tests)
mkdir -p gpurun_out; rm -f gpurun_out/stride.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "search_paths or odd_shapes" 2>&1 | tail -14 > gpurun_out/retry_tests.txt
cat gpurun_out/retry_tests.txt
for sr in ${SWEEP:-"16 16" "32 8" "32 9"}; do
  set -- $sr
  FWAV_UMMA_STRIDE=$1 FWAV_UMMA_RANK=$2 FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > gpurun_out/stride_run.out 2> gpurun_out/stride_run.err
  echo "== stride $1 rank $2" >> gpurun_out/stride.txt
  grep "fwav\]" gpurun_out/stride_run.err | tail -3 | cut -c1-200 >> gpurun_out/stride.txt
  cut -c1-220 gpurun_out/stride_run.out >> gpurun_out/stride.txt
done
cat gpurun_out/stride.txt
