#!/bin/bash
# round 2, call Y: collect pass against the per-stage budget for parked hits
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02y_hit_budget.txt
for b in 0 150 250 350 450 600 800 100000; do
  echo "== FWAV_UMMA_HIT_BUDGET=$b" >> $O/r02y_hit_budget.txt
  FWAV_UMMA_HIT_BUDGET=$b timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-170 >> $O/r02y_hit_budget.txt
done
cat $O/r02y_hit_budget.txt
