#!/bin/bash
# Round 2, GPU call J: half-precision accumulators in the collect pass -- accuracy probe, parity, timing
set +e
O=gpurun_out; mkdir -p $O
timeout 300 scripts/umma_f16acc_probe acc 300 > $O/r02j_f16acc_probe.json 2>&1; cat $O/r02j_f16acc_probe.json
FWAV_UMMA_ACC16=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "search or multi_batch or config2 or config1 or adversarial or end_to_end" > $O/r02j_tests.txt 2>&1
echo "tests with FWAV_UMMA_ACC16=1: rc=$? $(tail -1 $O/r02j_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02j_tests.txt | head
rm -f $O/r02j_timing.txt
for a in 0 1; do
  echo "== FWAV_UMMA_ACC16=$a (config 2)" >> $O/r02j_timing.txt
  FWAV_UMMA_ACC16=$a FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > $O/r02j_t.out 2> $O/r02j_t.err
  grep "fwav\]" $O/r02j_t.err | tail -3 | cut -c1-260 >> $O/r02j_timing.txt
  cut -c1-330 $O/r02j_t.out >> $O/r02j_timing.txt
done
cat $O/r02j_timing.txt
