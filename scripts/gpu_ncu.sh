#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/time_topk.py 0.12 umma 1 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk_umma -c 1 -o gpurun_out/prof_umma_v2 -f python scripts/time_topk.py 0.12 umma 1 > gpurun_out/ncu_umma.log 2>&1
tail -2 gpurun_out/ncu_umma.log; cat gpurun_out/ncu_plain.log | tail -1
