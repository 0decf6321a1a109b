#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dbg_modes.jsonl
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "umma or large_search" > gpurun_out/pytest_umma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_umma.log
tail -4 gpurun_out/pytest_umma.log
for d in 0 4 2 10; do FWAV_UMMA_DEBUG=$d timeout 120 python scripts/time_topk.py 0.25 umma 3 2>&1 | tail -1; done >> gpurun_out/dbg_modes.jsonl
timeout 200 python scripts/time_topk.py 1.0 umma 2 2>&1 | tail -1 >> gpurun_out/dbg_modes.jsonl
cat gpurun_out/dbg_modes.jsonl
