#!/bin/bash
# round 2: route choice by cost model -- config 2 whole, config 2 in 8 batches (what 8 ranks see), a 15-minute signal
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02_route.txt
run() { echo "== $1" >> $O/r02_route.txt; shift; env "$@" FWAV_UMMA_VERBOSE=1 timeout 300 python scripts/time_topk.py $SCALE umma 2 2> $O/r02_route.err | cut -c1-250 >> $O/r02_route.txt; grep "lack the room" $O/r02_route.err | sort | uniq -c | sort -rn | head -12 | cut -c1-230 >> $O/r02_route.txt; grep "second chance" $O/r02_route.err | tail -3 | cut -c1-200 >> $O/r02_route.txt; }
SCALE=1.0
run "config 2, one batch" X=1
run "config 2, batches of 62016 queries" FWAV_UMMA_BATCH=62016
run "config 2, batches of 62016 queries, fp16 accumulators forced" FWAV_UMMA_BATCH=62016 FWAV_UMMA_MODE=acc16
run "config 2, batches of 62016 queries, float32 hi*hi forced" FWAV_UMMA_BATCH=62016 FWAV_UMMA_MODE=hionly
SCALE=5.0
run "15 minutes (2.48 M queries x 9.9 M domains), cost model" X=1
run "15 minutes, full split forced" FWAV_UMMA_MODE=precise
run "15 minutes, fp16 accumulators forced" FWAV_UMMA_MODE=acc16
run "15 minutes, float32 hi*hi forced" FWAV_UMMA_MODE=hionly
cat $O/r02_route.txt
