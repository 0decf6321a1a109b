#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for e in 1 0; do
  if [ $e = 1 ]; then export FWAV_UMMA_THETA_FULL=1; else unset FWAV_UMMA_THETA_FULL; fi
  echo "== THETA_FULL=$e"; FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 2> $O/r02q2_t.err | cut -c1-330
  grep "fwav\]" $O/r02q2_t.err | tail -2 | cut -c1-220
done
unset FWAV_UMMA_THETA_FULL
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial or topk" > $O/r02q2_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02q2_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02q2_tests.txt | head -5
