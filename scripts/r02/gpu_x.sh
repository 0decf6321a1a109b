#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
L=$PWD/audio-compression_b200/fwav_b200/libfwav_b200_dbg.so
FWAV_LIB=$L FWAV_UMMA_DEBUG=512 timeout 200 python scripts/time_topk.py 1.0 umma 1 > $O/r02x.out 2> $O/r02x_hist.txt
cut -c1-200 $O/r02x.out; grep "rows with hits" $O/r02x_hist.txt | tail -1
