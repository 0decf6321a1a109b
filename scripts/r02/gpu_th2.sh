#!/bin/bash
# round 2: threshold pass through collect_hi_kernel<true> (FWAV_UMMA_THETA16=1, default) against scan_kernel (=0)
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02_theta16.txt
run() { echo "== $1" >> $O/r02_theta16.txt; shift; env "$@" FWAV_UMMA_VERBOSE=1 timeout 300 python scripts/time_topk.py $SCALE umma 2 2> $O/r02_theta16.err | cut -c1-250 >> $O/r02_theta16.txt; grep "lack the room" $O/r02_theta16.err | sort | uniq -c | sort -rn | head -4 | cut -c1-230 >> $O/r02_theta16.txt; grep "second chance\|to the exact list" $O/r02_theta16.err | tail -4 | cut -c1-200 >> $O/r02_theta16.txt; }
SCALE=1.0
run "config 2, THETA16=1" FWAV_UMMA_THETA16=1
run "config 2, THETA16=0" FWAV_UMMA_THETA16=0
run "config 2, THETA16=1 rank 7" FWAV_UMMA_THETA16=1 FWAV_UMMA_RANK=7
run "config 2, THETA16=1 rank 6" FWAV_UMMA_THETA16=1 FWAV_UMMA_RANK=6
SCALE=5.0
run "15 minutes, THETA16=1" FWAV_UMMA_THETA16=1
run "15 minutes, THETA16=0" FWAV_UMMA_THETA16=0
cat $O/r02_theta16.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial or topk" > $O/r02_theta16_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02_theta16_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02_theta16_tests.txt | head -5
