#!/bin/bash
# round 2, call S: search timing on both routes + the whole GPU suite, library as built
set +e
O=gpurun_out; mkdir -p $O
for acc in 1 0; do echo "== FWAV_UMMA_ACC16=$acc"; FWAV_UMMA_ACC16=$acc timeout 200 python scripts/time_topk.py 1.0 umma 3 2> $O/r02s_t.err | cut -c1-330; done
echo "== FWAV_UMMA_MODE=precise"; FWAV_UMMA_MODE=precise timeout 200 python scripts/time_topk.py 1.0 umma 2 2>> $O/r02s_t.err | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02s_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02s_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02s_tests.txt | head -5
