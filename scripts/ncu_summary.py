#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON: per kernel the metrics the roofline
discussion uses.  usage: ncu_summary.py report.ncu-rep out.json"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct", "lts__t_sector_op_write_hit_rate.pct",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].replace("<unnamed>::", "")}
    for i, h in enumerate(hdr):
        if h in KEYS and r[i] not in ("", "n/a"):
            d[h] = {"value": r[i], "unit": units[i]}
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
            d.setdefault("stall_per_issue", {})[h[len(STALL):-len("_per_issue_active.ratio")]] = float(r[i])
    tm = [h for h in hdr if "tensor" in h and "pct" in h]
    for h in tm:
        v = r[hdr.index(h)]
        if v not in ("", "n/a", "0"):
            d.setdefault("tensor_metrics", {})[h] = v
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
for d in res:
    t = d.get("gpu__time_duration.sum", {}).get("value")
    print(d["kernel"][:60], t, d.get("dram__bytes_read.sum", {}).get("value"), d.get("dram__bytes_write.sum", {}).get("value"),
          d.get("tensor_metrics"))
