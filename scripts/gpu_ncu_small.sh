#!/bin/bash
# ncu metrics of the HBM-bound kernels (domains, embed, affine, decode) in one bench run
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --decode-scale 0.5 > gpurun_out/small_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"half_sums|domains_from|embed_static|affine_kernel|decode_iter" -s 8 -c 10 --csv --log-file gpurun_out/small_kernels.csv python bench.py --steps 1 --warmup 3 --no-cpu --decode-scale 0.5 > gpurun_out/small_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/small_kernels.csv")) if len(r)>10 and r[0].isdigit()]
d=collections.OrderedDict()
for r in rows:
    key=(r[0], r[4].split('(')[0][-44:])
    d.setdefault(key,{})[r[-3]]=(r[-1],r[-2])
for k,v in d.items():
    print(k[1], {m.split('.')[0][-28:]+'.'+m.split('.')[-1][:6]:x[0]+x[1] for m,x in v.items()})
PY
