#!/bin/bash
# round 2, call U: tcgen05.ld drain rates, float32 and pack::16b, by number of loading warps
mkdir -p gpurun_out; rm -f gpurun_out/r02u_ld_microbench.jsonl
for v in ld_1w ld_4w ld_8w ld_16w ld_1w_pack16 ld_4w_pack16 ld_8w_pack16 ld_16w_pack16 ld_8w_max ld_16w_max ld_8w_pack16_max ld_16w_pack16_max; do
  timeout 60 scripts/umma_microbench $v >> gpurun_out/r02u_ld_microbench.jsonl 2>&1
done
cat gpurun_out/r02u_ld_microbench.jsonl
