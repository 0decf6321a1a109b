#!/bin/bash
# End-of-round pass on one GPU: parity tests, smoke, bench, ncu launch list (reference arm: scripts/gpu_full.sh).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; head -c 1200 gpurun_out/bench.json; echo
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > gpurun_out/ncu_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
