#!/bin/bash
# Full single-GPU pass: parity tests, smoke, bench (ours + reference arm), ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?" >> gpurun_out/bench_reference.err
head -c 600 gpurun_out/bench_reference.json; echo
if [ "$1" == "ncu" ]; then
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
fi
