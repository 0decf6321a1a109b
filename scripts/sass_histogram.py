#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), the evidence that the search is tcgen05 /
TMEM / bulk-copy code: usage: sass_histogram.py libfwav_b200.so out.txt"""
import collections, re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "UTCATOM", "FMNMX3", "DFMA", "FFMA", "STG", "LDG", "MULTIMEM", "RED", "ATOM")
with open(out, "w") as f:
    f.write("# cuobjdump -sass %s : instructions per kernel, then the opcodes that matter (prefix match)\n" % lib.split("/")[-1])
    for k, c in hist.items():
        tot = sum(c.values())
        sel = {p: sum(v for o, v in c.items() if o.startswith(p)) for p in KEY}
        f.write("%-110s total %6d  " % (k[:110], tot) + " ".join("%s=%d" % (p, v) for p, v in sel.items() if v) + "\n")
    f.write("\n# full histograms: the dominant kernel (collect pass, fp16 accumulators), its float32-accumulator form, the verification pass\n")
    for k, c in hist.items():
        if "collect_hi_kernel<false>" in k or "scan_kernel<2, true, 1" in k or k.startswith("finalize_kernel"):
            f.write(k + "\n" + "\n".join("  %-28s %d" % (o, v) for o, v in c.most_common()) + "\n")
print("kernels:", len(hist))
