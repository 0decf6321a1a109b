#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dbg_modes.jsonl
for v in ts ss; do for d in 2 10 26; do FWAV_UMMA_VARIANT=$v FWAV_UMMA_DEBUG=$d timeout 120 python scripts/time_topk.py 0.25 umma 3 2>&1 | tail -1 | sed "s/^/$v /"; done; done >> gpurun_out/dbg_modes.jsonl
cat gpurun_out/dbg_modes.jsonl
