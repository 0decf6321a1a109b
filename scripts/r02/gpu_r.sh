#!/bin/bash
# round 2, call R: search timing, fp16-accumulator route and float32 route, library as built
set +e
O=gpurun_out; mkdir -p $O
for acc in 1 0; do echo "== FWAV_UMMA_ACC16=$acc"; FWAV_UMMA_ACC16=$acc timeout 200 python scripts/time_topk.py 1.0 umma 3 2> $O/r02r_t.err | cut -c1-330; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or multi_batch or config2 or adversarial or topk or compact or odd" > $O/r02r_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02r_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02r_tests.txt | head -5
