#!/bin/bash
# 8-GPU runs: config 2 sharded by ranges (strong scaling, incl. the sharded decode leg), then config 3 (1 h / 48 kHz)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?" >> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err; head -c 600 gpurun_out/bench_n$N.json; echo
if [ "$2" == "c3" ]; then
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload c3 --steps 2 --warmup 3 --no-decode > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "exit $?" >> gpurun_out/bench_c3_n$N.err
tail -3 gpurun_out/bench_c3_n$N.err; head -c 1200 gpurun_out/bench_c3_n$N.json; echo
fi
free -g | head -2
