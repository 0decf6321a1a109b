#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/n2b_bench.json 2> $O/n2b_bench.err
echo "bench exit $?"; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/n2b_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], {k: round(v["ms"], 3) for k, v in d["kernels"].items()})
print({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ("ms_per_iter", "value")}) for k, v in (d.get("decode") or {}).items()})
PY
