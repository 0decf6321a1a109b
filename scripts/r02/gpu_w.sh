#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02w_mma_latency.txt
for n in 1 1 2 4 8 16; do timeout 30 scripts/umma_f16acc_probe 0 $n | head -1 >> gpurun_out/r02w_mma_latency.txt; done
for n in 1 2 4; do timeout 30 scripts/umma_f16acc_probe 1 $n | head -1 >> gpurun_out/r02w_mma_latency.txt; done
cat gpurun_out/r02w_mma_latency.txt
