#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 -k "umma or large_search" > gpurun_out/pytest_umma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_umma.log
tail -25 gpurun_out/pytest_umma.log
timeout 300 python bench.py --steps 3 --warmup 3 --search umma --no-cpu > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err
echo "bench exit $?" >> gpurun_out/bench_umma.err
tail -3 gpurun_out/bench_umma.err; head -c 2500 gpurun_out/bench_umma.json
