#!/bin/bash
# one `ncu --set full` capture of the search kernels (scan_kernel<THETA>, <COLLECT>, finalize) on a quarter-scale config 2
mkdir -p gpurun_out
timeout 200 python scripts/time_topk.py ${1:-0.25} umma 0 > gpurun_out/ncu_full_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|finalize_kernel" -c 3 -f -o gpurun_out/prof_topk_r01 python scripts/time_topk.py ${1:-0.25} umma 0 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
