#!/bin/bash
# 32-byte row requests (finalize, affine, table stores), fp16 second chance only with room.  Parity, timing, ncu.
set +e
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q -x --timeout=600 > $O/af_pytest.log 2>&1
echo "pytest exit $?" >> $O/af_pytest.log; tail -3 $O/af_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/af_bench.json 2> $O/af_bench.err; echo "bench exit $?"; tail -1 $O/af_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/af_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], "e2e", d["e2e"]["ms_per_step"])
PY
timeout 300 python scripts/time_topk.py 5.0 umma 1 2>/dev/null | cut -c1-330 | tee $O/af_time5.txt
timeout 600 ncu --set full --clock-control none -k regex:"tables_from_halves|affine_kernel|finalize_kernel" -c 3 -f -o $O/af_small python bench.py --steps 1 --warmup 0 --no-cpu --no-decode > $O/af_ncu.log 2>&1
tail -1 $O/af_ncu.log | cut -c1-200
