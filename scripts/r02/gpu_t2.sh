#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02t2_rank_sweep.txt
for sr in "32 8" "64 4" "64 5" "48 6" "48 5" "40 7" "32 7" "32 9"; do
  set -- $sr
  echo "== FWAV_UMMA_STRIDE=$1 FWAV_UMMA_RANK=$2 (expected candidates per query: $(( $1 * $2 )))" >> $O/r02t2_rank_sweep.txt
  FWAV_UMMA_STRIDE=$1 FWAV_UMMA_RANK=$2 FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 2 2> $O/r02t2.err | cut -c1-260 >> $O/r02t2_rank_sweep.txt
  grep "second chance\|lack the room" $O/r02t2.err | tail -2 | cut -c1-220 >> $O/r02t2_rank_sweep.txt
done
cat $O/r02t2_rank_sweep.txt
