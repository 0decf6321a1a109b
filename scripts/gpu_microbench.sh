#!/bin/bash
# runs every variant of scripts/umma_microbench in its own process
mkdir -p gpurun_out; out=gpurun_out/umma_microbench.jsonl; rm -f $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader >> $out
for v in $(scripts/umma_microbench list); do
  timeout 60 scripts/umma_microbench $v >> $out 2>&1 || echo "{\"variant\": \"$v\", \"exit\": $?}" >> $out
done
cat $out
