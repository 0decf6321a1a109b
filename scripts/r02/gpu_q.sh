#!/bin/bash
# round 2, call Q: timing of the search (config 2) + the search parity subset with the library as built
set +e
O=gpurun_out; mkdir -p $O
for r in 1 2; do FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 2> $O/r02q_t.err | cut -c1-330; done
grep "fwav\]" $O/r02q_t.err | tail -1 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial or topk" > $O/r02q_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02q_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02q_tests.txt | head -5
