#!/bin/bash
# Fused tables (chain half sums + rows embedded in registers) and the shared table position of collect_hi_kernel:
# parity of the new path, A/B timing of the position knob, one ncu --set full of the collect pass with it on.
set +e
O=gpurun_out; mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "tables or domains or embed or search_paths or multi_batch or large_search or config2_full" > $O/aa_pytest.log 2>&1
echo "pytest exit $?" >> $O/aa_pytest.log; tail -4 $O/aa_pytest.log
for f in 1 0 1 0; do FWAV_UMMA_FRONT=$f timeout 200 python scripts/time_topk.py 1.0 umma 4 2>/dev/null | cut -c1-400; done | tee $O/aa_front.txt
FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 0 2> $O/aa_verbose.txt > /dev/null; cat $O/aa_verbose.txt | cut -c1-250
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/aa_bench.json 2> $O/aa_bench.err; echo "bench exit $?"; tail -2 $O/aa_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/aa_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["kernels"], d["roofline"]["search_phases_ms"], d["e2e"]["ms_per_step"])
PY
timeout 600 ncu --set full --clock-control none -k regex:collect_hi_kernel -s 1 -c 1 -f -o $O/aa_collect_front python scripts/time_topk.py 1.0 umma 0 > $O/aa_ncu.log 2>&1
tail -2 $O/aa_ncu.log | cut -c1-200; ls -la $O/*.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"half_sums_chain|tables_from_halves" -c 2 -f -o $O/aa_tables python scripts/time_topk.py 1.0 umma 0 > $O/aa_ncu2.log 2>&1
tail -2 $O/aa_ncu2.log | cut -c1-200
