#!/bin/bash
# round 2, call Z: collect_hi_kernel tuning variants (scripts/r02/build_variants.sh), config 2, same box
set +e
O=gpurun_out; mkdir -p $O; rm -f $O/r02z_variants.txt
for v in head r1c1w1 r0c1w1 r1c0w1 r0c0w1 r1c1w0 r0c0w0 head; do
  echo "== $v" >> $O/r02z_variants.txt
  FWAV_LIB=$PWD/audio-compression_b200/fwav_b200/variants/libfwav_b200_$v.so timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-150 >> $O/r02z_variants.txt
done
cat $O/r02z_variants.txt
