#!/usr/bin/env python
"""Per-phase instruction / stall-sample breakdown of the tile kernel from an ncu report: phases are
delimited by the '// ---- N.' markers and the helper-function headers in srb_kernels_tile.cuh.
The report must have been taken from the same source revision.
    python tools/phase_breakdown.py gpurun_out/prof.ncu-rep [pixels_per_launch]"""
import csv, re, subprocess, sys
csv.field_size_limit(10 ** 9)
SRC = sys.argv[3] if len(sys.argv) > 3 else "super-resolution_b200/csrc/srb_kernels_tile.cuh"
rep = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else 2048 * 2048 * 3
marks = []
for i, line in enumerate(open(SRC), 1):
    m = re.match(r"\s*// ---- (\w+)\.", line)
    if m:
        marks.append((i, "phase " + m.group(1)))
    m = re.match(r"__device__ __forceinline__ \w+[\s\*&]+(\w+)\(", line)
    if m:
        marks.append((i - 1, "fn " + m.group(1)))
    if line.startswith("k_tile("):
        marks.append((i - 2, "k_tile prologue"))
    if "cost partial sums" in line:
        marks.append((i, "cost reduce"))
marks.sort()
def phase_of(ln):
    name = "other"
    for start, n in marks:
        if ln >= start:
            name = n
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, acc, ti, ts = None, {}, 0, 0
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln, smp, ins = int(r[0]), int(r[4]), int(r[7])
    except ValueError:
        continue
    k = phase_of(ln) if cur == "srb_kernels_tile.cuh" else "other files (%s)" % cur
    a = acc.setdefault(k, [0, 0])
    a[0] += ins
    a[1] += smp
    ti += ins
    ts += smp
print("%-28s %8s %8s %10s" % ("phase", "inst %", "stall %", "inst/px"))
for k, (i, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %7.1f%% %7.1f%% %10.1f" % (k, 100.0 * i / ti, 100.0 * s / max(ts, 1), i * 32.0 / npx))
print("%-28s %8s %8s %10.1f" % ("total", "", "", ti * 32.0 / npx))
