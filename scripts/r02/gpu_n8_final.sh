#!/bin/bash
# N GPUs, config 2 only (no extras): the end-of-round kernels on the scaling curve.
set +e
N=${1:-8}
O=gpurun_out; mkdir -p $O
FWAV_BENCH_EXTRAS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $O/n${N}f_bench.json 2> $O/n${N}f_bench.err
echo "bench exit $?"; N=$N python - <<PY
import json, os
N = os.environ["N"]
d = json.loads([l for l in open(f"gpurun_out/n{N}f_bench.json") if l.startswith("{")][-1])
print(d["details"]["per_rank"]); print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"], {k: round(v["ms"], 3) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], d["clocks"])
print({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ("ms_per_iter", "value")}) for k, v in (d.get("decode") or {}).items()})
PY
