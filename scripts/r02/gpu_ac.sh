#!/bin/bash
# Half-length DCT chains in the static embedding; L2 evict_last hint on the table tiles of the collect pass (knob).
set +e
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "tables or domains or embed or affine or search_paths or fixed_mode or query_mode or config1 or golden or reference_own" > $O/ac_pytest.log 2>&1
echo "pytest exit $?" >> $O/ac_pytest.log; tail -4 $O/ac_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/ac_bench.json 2> $O/ac_bench.err; echo "bench exit $?"; tail -1 $O/ac_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/ac_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], "e2e", d["e2e"]["ms_per_step"])
PY
for k in 0 1; do FWAV_UMMA_L2KEEP=$k timeout 200 python scripts/time_topk.py 1.0 umma 3 2>/dev/null | cut -c1-330; done | tee $O/ac_l2keep.txt
FWAV_UMMA_L2KEEP=1 timeout 600 ncu --set full --clock-control none -k regex:collect_hi_kernel -s 1 -c 1 -f -o $O/ac_collect_l2keep python scripts/time_topk.py 1.0 umma 0 > $O/ac_ncu.log 2>&1
tail -1 $O/ac_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"half_sums_chain|tables_from_halves|affine_kernel" -c 3 -f -o $O/ac_small python bench.py --steps 1 --warmup 0 --no-cpu --no-decode > $O/ac_ncu2.log 2>&1
tail -1 $O/ac_ncu2.log | cut -c1-200
