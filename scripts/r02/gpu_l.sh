#!/bin/bash
# Round 2, GPU call L: four issuing threads in the hi*hi-only collect pass (collect_hi_kernel), with float32 and
# half-precision accumulators -- parity, then timing of the four combinations on config 2
set +e
O=gpurun_out; mkdir -p $O
for a in 0 1; do
  FWAV_UMMA_ISSUERS=4 FWAV_UMMA_ACC16=$a timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial" > $O/r02l_tests_$a.txt 2>&1
  echo "tests ISSUERS=4 ACC16=$a: rc=$? $(tail -1 $O/r02l_tests_$a.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02l_tests_$a.txt | head -5
done
rm -f $O/r02l_timing.txt
for i in 2 4; do for a in 0 1; do
  echo "== FWAV_UMMA_ISSUERS=$i FWAV_UMMA_ACC16=$a (config 2)" >> $O/r02l_timing.txt
  FWAV_UMMA_ISSUERS=$i FWAV_UMMA_ACC16=$a FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > $O/r02l_t.out 2> $O/r02l_t.err
  grep "fwav\]" $O/r02l_t.err | tail -2 | cut -c1-200 >> $O/r02l_timing.txt
  cut -c1-330 $O/r02l_t.out >> $O/r02l_timing.txt
done; done
cat $O/r02l_timing.txt
