#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report:  python tools/ncu_hot_lines.py rep.ncu-rep [N]
(needs the kernel compiled with -lineinfo and captured with --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; hdr = None; data = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; i_s = r.index("# Samples"); i_i = r.index("Instructions Executed"); continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0].isdigit():
        try: data.append((cur, int(r[0]), r[1], int(r[i_s] or 0), int(r[i_i] or 0)))
        except ValueError: pass
ts = sum(d[3] for d in data) or 1; ti = sum(d[4] for d in data) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-16s %4d %6.2f%% smp %6.2f%% inst  %s" % (d[0], d[1], 100.0 * d[3] / ts, 100.0 * d[4] / ti, d[2].strip()[:90]))
