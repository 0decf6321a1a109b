#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?" >> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err; head -c 3500 gpurun_out/bench_n$N.json
