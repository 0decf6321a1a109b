#!/bin/bash
# Full GPU suite with the fused tables, the pipelined affine kernel and the lowered threshold for short queries;
# A/B of the affine forms; bench; ncu of the small kernels.
set +e
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q --timeout=600 > $O/ab_pytest.log 2>&1
echo "pytest exit $?" >> $O/ab_pytest.log; tail -5 $O/ab_pytest.log
for v in 1 0; do
  FWAV_AFFINE_PIPE=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/ab_bench_pipe$v.json 2> $O/ab_bench_pipe$v.err; echo "bench pipe=$v exit $?"; tail -1 $O/ab_bench_pipe$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_bench_pipe$v.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], "e2e", d["e2e"]["ms_per_step"])
PY
done
FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 0 2> $O/ab_verbose.txt > /dev/null; grep fwav $O/ab_verbose.txt | cut -c1-250
timeout 600 ncu --set full --clock-control none -k regex:"half_sums_chain|tables_from_halves|affine_kernel" -c 3 -f -o $O/ab_small python bench.py --steps 1 --warmup 0 --no-cpu --no-decode > $O/ab_ncu.log 2>&1
tail -2 $O/ab_ncu.log | cut -c1-200
