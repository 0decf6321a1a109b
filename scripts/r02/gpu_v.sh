#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/mbar_microbench > gpurun_out/r02v_mbar_microbench.jsonl 2>&1; cat gpurun_out/r02v_mbar_microbench.jsonl
