#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_fused_decode.py > $O/r02h_fused.txt 2>&1
echo "fused decode test rc=$?"; grep -vE "^\*|OMP_NUM" $O/r02h_fused.txt | tail -25 | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02h_bench_n2.json 2> $O/r02h_bench_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step']); print('decode', d['decode'])
PY
