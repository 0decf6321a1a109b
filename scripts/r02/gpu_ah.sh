#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "search or multi_batch or affine or config2_full or aligned" > $O/ah_pytest.log 2>&1
echo "pytest exit $?" >> $O/ah_pytest.log; tail -2 $O/ah_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/ah_bench.json 2> $O/ah_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/ah_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], "e2e", d["e2e"]["ms_per_step"])
PY
