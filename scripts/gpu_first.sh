#!/bin/bash
# First GPU pass: parity tests, smoke, a short bench.  Logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q -s --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --search ffma > gpurun_out/bench_ffma.json 2> gpurun_out/bench_ffma.err
echo "bench exit $?" >> gpurun_out/bench_ffma.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_ffma.json | head -c 3000
