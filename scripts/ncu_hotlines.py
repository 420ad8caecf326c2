#!/usr/bin/env python
"""Per-CUDA-line summary of one kernel of an ncu report captured with --set full --import-source on:

    python scripts/ncu_hotlines.py report.ncu-rep <kernel regex> [top N]

Prints, per source line, the share of warp-stall samples, the share of executed warp instructions and the average
number of active threads per executed instruction (lane efficiency), sorted by instruction share.  The regex matches the kernel's base name; the first matching launch
of the report is taken."""
import collections
import csv
import io
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + regex, "-c", "1"],
                     capture_output=True, text=True).stdout
cur, hdr, kernel, agg = None, None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        kernel = r[1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].strip():          # a CUDA line (its SASS rows follow with an empty line number)
        try:
            agg[(cur, int(r[0]))] = (int(r[4]), int(r[7]), int(r[8]), r[1].strip())
        except ValueError:
            pass
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("# %s" % kernel)
print("# warp-stall samples %d, warp instructions %d, threads per instruction %.1f" % (ts, ti, sum(a[2] for a in agg.values())/ti))
byfile = collections.defaultdict(lambda: [0, 0])
for (f, _), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print("# %-32s %5.1f %% of samples, %5.1f %% of instructions" % (f, 100*a[0]/ts, 100*a[1]/ti))
print("# %samples %instructions threads/inst file:line source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f %5.1f %5.1f  %s:%d  %s" % (100*a[0]/ts, 100*a[1]/ti, a[2]/max(a[1], 1), f, l, a[3][:130]))
