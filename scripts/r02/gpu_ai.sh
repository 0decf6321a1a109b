#!/bin/bash
# fp16 second chance with >= 4 096 entries per query (larger buffers where the failures are many); 2 MB ring chunks.
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=500 -k "search or multi_batch or config2_full or adversarial or config4 or config1 or reference_own or pinned or host" > $O/ai_pytest.log 2>&1
echo "pytest exit $?" >> $O/ai_pytest.log; tail -2 $O/ai_pytest.log
FWAV_UMMA_VERBOSE=1 timeout 300 python scripts/time_topk.py 5.0 umma 1 2> $O/ai_verbose5.txt | cut -c1-330; grep "second chance\|list kernel" $O/ai_verbose5.txt | tail -4 | cut -c1-250
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode > $O/ai_bench.json 2> $O/ai_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/ai_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["roofline"]["search_phases_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["c_abi_pinned_ms"])
PY
