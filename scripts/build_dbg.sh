#!/bin/bash
# debug build of the library (profiling knobs compiled in: FWAV_UMMA_DEBUG, trace stamps); used through FWAV_LIB=...
set -e
cd "$(dirname "$0")/../audio-compression_b200/csrc"
mkdir -p build_dbg
for f in api prestep domains embed topk_ffma topk_umma affine decode; do
  if [ ! -f build_dbg/$f.o ] || [ $f.cu -nt build_dbg/$f.o ] || [ common.cuh -nt build_dbg/$f.o ]; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
      --expt-relaxed-constexpr -DFWAV_DEBUG_KNOBS -c $f.cu -o build_dbg/$f.o
  fi
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../fwav_b200/libfwav_b200_dbg.so build_dbg/*.o -lcudart
echo built ../fwav_b200/libfwav_b200_dbg.so
