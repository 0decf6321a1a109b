#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
FWAV_UMMA_ISSUERS=44 FWAV_UMMA_ACC16=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial" > $O/r02m_tests.txt 2>&1
echo "tests ISSUERS=44 ACC16=1: rc=$? $(tail -1 $O/r02m_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02m_tests.txt | head -5
rm -f $O/r02m_timing.txt
for i in 4 44; do
  echo "== FWAV_UMMA_ISSUERS=$i FWAV_UMMA_ACC16=1 (config 2)" >> $O/r02m_timing.txt
  FWAV_UMMA_ISSUERS=$i FWAV_UMMA_ACC16=1 FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > $O/r02m_t.out 2> $O/r02m_t.err
  grep "fwav\]" $O/r02m_t.err | tail -2 | cut -c1-200 >> $O/r02m_timing.txt
  cut -c1-330 $O/r02m_t.out >> $O/r02m_timing.txt
done
cat $O/r02m_timing.txt
