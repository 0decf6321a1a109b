#!/bin/bash
mkdir -p gpurun_out
for d in 0 4 1 2; do FWAV_UMMA_DEBUG=$d timeout 120 python scripts/time_topk.py 0.25 umma 3; done > gpurun_out/dbg_modes.jsonl 2>&1
timeout 120 python scripts/time_topk.py 0.25 ffma 2 >> gpurun_out/dbg_modes.jsonl 2>&1
cat gpurun_out/dbg_modes.jsonl
timeout 120 python scripts/time_topk.py 0.06 umma 1 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_umma -c 1 -o gpurun_out/prof_umma_v1 -f python scripts/time_topk.py 0.06 umma 1 > gpurun_out/ncu_umma.log 2>&1
tail -3 gpurun_out/ncu_umma.log
