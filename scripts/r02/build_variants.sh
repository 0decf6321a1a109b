#!/bin/bash
# builds libfwav_b200 variants of the collect_hi_kernel tuning switches into audio-compression_b200/fwav_b200/variants/
set -e
cd "$(dirname "$0")/../../audio-compression_b200/csrc"
mkdir -p ../fwav_b200/variants build_var
for v in "r1c1w1" "r0c1w1" "r1c0w1" "r0c0w1" "r1c1w0" "r0c0w0"; do
  R=${v:1:1}; C=${v:3:1}; W=${v:5:1}
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DFWAV_HI_ROTATE=$R -DFWAV_HI_CALL=$C -DFWAV_HI_ROWS2=$W -c topk_umma.cu -o build_var/topk_umma_$v.o &
done
wait
for v in "r1c1w1" "r0c1w1" "r1c0w1" "r0c0w1" "r1c1w0" "r0c0w0"; do
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../fwav_b200/variants/libfwav_b200_$v.so \
     build/api.o build/prestep.o build/domains.o build/embed.o build/topk_ffma.o build_var/topk_umma_$v.o build/affine.o build/decode.o -lcudart
done
ls -la ../fwav_b200/variants/
