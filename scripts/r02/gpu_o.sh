#!/bin/bash
# round 2, call O: pipe microbenchmark (which pipe / rate for the packed max family) + threshold-in-accumulator probe
mkdir -p gpurun_out
timeout 120 scripts/pipe_microbench > gpurun_out/r02o_pipe_microbench.jsonl 2>&1
timeout 300 scripts/umma_f16acc_probe bias 300 > gpurun_out/r02o_bias_probe.json 2>&1
cat gpurun_out/r02o_pipe_microbench.jsonl gpurun_out/r02o_bias_probe.json
