#!/bin/bash
# round 2, call P: clock64 trace of one CTA of collect_hi_kernel (debug build, FWAV_UMMA_DEBUG=64)
set +e
O=gpurun_out; mkdir -p $O
L=$PWD/audio-compression_b200/fwav_b200/libfwav_b200_dbg.so
FWAV_LIB=$L FWAV_UMMA_DEBUG=64 timeout 200 python scripts/time_topk.py 1.0 umma 1 > $O/r02p_trace.out 2> $O/r02p_trace.err
cut -c1-300 $O/r02p_trace.out; tail -70 $O/r02p_trace.err
