#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02k_timing.txt
L=$PWD/audio-compression_b200/fwav_b200/libfwav_b200_dbg.so
for a in 0 1; do for d in 0 4 8; do
  echo "== debug build, FWAV_UMMA_ACC16=$a FWAV_UMMA_DEBUG=$d (4: thresholds at +inf, 8: MMA free-running)" >> $O/r02k_timing.txt
  FWAV_LIB=$L FWAV_UMMA_ACC16=$a FWAV_UMMA_DEBUG=$d timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-200 >> $O/r02k_timing.txt
done; done
cat $O/r02k_timing.txt
