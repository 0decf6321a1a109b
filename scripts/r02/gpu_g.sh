#!/bin/bash
# Round 2, GPU call G (8 GPUs): the driver's --gpus 8 command, extras included (c3, c5 from c3, c4 at 1/4 length)
set +e
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02g_smi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r02g_bench_n8.json 2> $O/r02g_bench_n8.err
echo "bench n8 rc=$?"; tail -3 $O/r02g_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_n8.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value']); print('e2e', d['e2e']); print('kern', {k:v.get('ms') for k,v in d['kernels'].items()}); print('search', d['roofline'].get('search_phases_ms')); print('decode', d['decode']); print('clocks', d['clocks']); print('extra', json.dumps(d['extra'])[:4000])
PY
