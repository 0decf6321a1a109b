#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a small JSON: per kernel
the number of launches, the summed duration and its share.  usage: ncu_launch_summary.py launches.csv out.json"""
import collections, csv, json, re, sys
src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4]).replace("<unnamed>::", "")
    name = name.split("(")[0]
    agg.setdefault(name, []).append(float(r[-1].replace(",", "")))
total = sum(sum(v) for v in agg.values())
doc = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 1 --warmup 3 --no-cpu --no-decode",
       "unit": "ns", "total_ns": total,
       "kernels": {n: {"launches": len(v), "total_ns": sum(v), "share": sum(v) / total}
                   for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps({n: round(k["share"], 4) for n, k in list(doc["kernels"].items())[:6]}))
