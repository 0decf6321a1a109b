#!/bin/bash
# Round 2, GPU call F (2 GPUs): new search tests (error bounds, adversarial table), bench under torchrun at N=2 with
# shortened extras, sharded decode timings
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "adversarial or query_mode or fixed_mode or search or multi_batch or compact or config" > $O/r02f_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02f_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02f_tests.txt | head
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "adversarial or query_mode or fixed_mode" 2>&1 | grep -E "adversarial|query_mode=range|fixed mode" | head
timeout 200 python scripts/time_topk.py 1.0 umma 3 2>/dev/null | cut -c1-330
FWAV_BENCH_EXTRAS=1 FWAV_BENCH_EXTRAS_SCALE=0.1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02f_bench_n2.json 2> $O/r02f_bench_n2.err
echo "bench n2 rc=$?"; tail -3 $O/r02f_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_n2.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value']); print('e2e', d['e2e']); print('kern', {k:v.get('ms') for k,v in d['kernels'].items()}); print('decode', d['decode']); print('extra', json.dumps(d['extra'])[:3000])
PY
