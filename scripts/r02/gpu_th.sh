#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
L=$PWD/audio-compression_b200/fwav_b200/libfwav_b200_dbg.so
for d in 0 4 2; do echo "== debug build FWAV_UMMA_DEBUG=$d"; FWAV_LIB=$L FWAV_UMMA_DEBUG=$d timeout 200 python scripts/time_topk.py 1.0 umma 1 2>/dev/null | cut -c1-200; done
