#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "umma or large_search" > gpurun_out/pytest_umma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_umma.log
tail -4 gpurun_out/pytest_umma.log
for d in 0 4 2; do FWAV_UMMA_DEBUG=$d timeout 120 python scripts/time_topk.py 0.25 umma 3; done > gpurun_out/dbg_modes.jsonl 2>&1
timeout 200 python scripts/time_topk.py 1.0 umma 2 >> gpurun_out/dbg_modes.jsonl 2>&1
FWAV_UMMA_VARIANT=ss timeout 200 python scripts/time_topk.py 1.0 umma 2 >> gpurun_out/dbg_modes.jsonl 2>&1
cat gpurun_out/dbg_modes.jsonl
