#!/bin/bash
# BASELINE.json config 4 at full length on 8 GPUs: 30 min / 48 kHz, tile 1024 (range_size 4, domain_step 1), top-K 64
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload c4 --steps 1 --warmup 1 --no-decode > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
echo "exit $?" >> gpurun_out/bench_c4_n$N.err
tail -3 gpurun_out/bench_c4_n$N.err; head -c 1500 gpurun_out/bench_c4_n$N.json; echo
