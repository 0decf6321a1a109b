#!/bin/bash
# `ncu --set full` of the collect pass (the dominant kernel) at full config-2 scale: DRAM traffic per launch etc.
mkdir -p gpurun_out
timeout 200 python scripts/time_topk.py 1.0 umma 0 > gpurun_out/ncu_collect_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_collect_full_r01 python scripts/time_topk.py 1.0 umma 0 > gpurun_out/ncu_collect.log 2>&1
tail -2 gpurun_out/ncu_collect.log; ls -la gpurun_out/prof_collect_full_r01.ncu-rep
