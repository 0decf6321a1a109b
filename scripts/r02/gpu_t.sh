#!/bin/bash
# round 2, call T: how the collect pass responds to the number of candidates per query (threshold sample stride x rank)
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02t_rank_sweep.txt
for sr in "32 12" "32 8" "32 6" "32 4" "16 12" "16 8" "16 6" "16 4" "8 8"; do
  set -- $sr
  echo "== FWAV_UMMA_STRIDE=$1 FWAV_UMMA_RANK=$2 (expected candidates per query: $(( $1 * $2 )))" >> $O/r02t_rank_sweep.txt
  FWAV_UMMA_STRIDE=$1 FWAV_UMMA_RANK=$2 FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 2 2> $O/r02t.err | cut -c1-260 >> $O/r02t_rank_sweep.txt
  grep "second chance\|lack the room" $O/r02t.err | tail -2 | cut -c1-220 >> $O/r02t_rank_sweep.txt
done
cat $O/r02t_rank_sweep.txt
