#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 1 2> $O/ag_verbose_topk.txt | cut -c1-300; grep fwav $O/ag_verbose_topk.txt | cut -c1-250
FWAV_UMMA_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > $O/ag_bench.json 2> $O/ag_bench.err; grep fwav $O/ag_bench.err | tail -8 | cut -c1-250
