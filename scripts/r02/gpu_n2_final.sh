#!/bin/bash
# Two GPUs: fused decode + broadcast test, bench with the extras (c3 / c5 / c4 shapes at 1/4 length) through the
# sharded driver, with the end-of-round kernels.
set +e
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_fused_decode.py > $O/n2_fused_decode.log 2>&1; echo "fused decode exit $?"; tail -3 $O/n2_fused_decode.log | cut -c1-300
FWAV_BENCH_EXTRAS=1 FWAV_BENCH_EXTRAS_SCALE=0.25 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/n2_bench.json 2> $O/n2_bench.err
echo "bench exit $?"; tail -3 $O/n2_bench.err | cut -c1-300; head -c 600 $O/n2_bench.json; echo
