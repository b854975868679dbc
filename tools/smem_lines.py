#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts / bank-conflict excess and stall reasons from an ncu report
(read here, no GPU needed).   python tools/smem_lines.py gpurun_out/prof.ncu-rep [top]"""
import csv, subprocess, sys
from collections import defaultdict
csv.field_size_limit(10 ** 9)
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
acc = defaultdict(lambda: defaultdict(float))
src = {}
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur, ln)
    src[key] = r[1].strip()[:90]
    for i, h in enumerate(hdr):
        if i < 4:
            continue
        try:
            acc[key][h] += float(r[i])
        except ValueError:
            pass
tot = defaultdict(float)
for k, d in acc.items():
    for h, v in d.items():
        tot[h] += v
W, X, I = "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "Instructions Executed"
print("total: instructions %.0f, smem wavefronts %.0f (excess %.0f), samples %.0f" %
      (tot[I], tot[W], tot[X], tot["# Samples"]))
print("\n-- by shared-memory wavefronts")
for k, d in sorted(acc.items(), key=lambda kv: -kv[1][W])[:top]:
    print("%5.1f%% wf  %5.1f%% of excess  %-26s %s" % (100 * d[W] / tot[W], 100 * d[X] / max(tot[X], 1),
                                                     "%s:%d" % k, src[k]))
print("\n-- stall reasons (all samples), share of total")
reasons = [h for h in tot if h.startswith("stall_") and "Not Issued" not in h]
s = sum(tot[h] for h in reasons)
for h in sorted(reasons, key=lambda h: -tot[h])[:10]:
    print("  %-24s %5.1f%%" % (h, 100 * tot[h] / s))
print("\n-- by stall samples")
for k, d in sorted(acc.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    rs = sorted(((d[h], h) for h in reasons), reverse=True)[:2]
    print("%5.1f%% smp  %-26s %-30s %s" % (100 * d["# Samples"] / tot["# Samples"], "%s:%d" % k,
                                        ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(d["# Samples"], 1)) for v, h in rs), src[k][:60]))
