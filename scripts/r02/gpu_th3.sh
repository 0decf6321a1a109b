#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for r in 8 7; do echo "== 15 minutes, rank $r"; FWAV_UMMA_RANK=$r timeout 300 python scripts/time_topk.py 5.0 umma 2 2>/dev/null | cut -c1-250; done
for r in 8 7; do echo "== config 2 x 0.25 (45 s), rank $r"; FWAV_UMMA_RANK=$r timeout 300 python scripts/time_topk.py 0.25 umma 3 2>/dev/null | cut -c1-250; done
