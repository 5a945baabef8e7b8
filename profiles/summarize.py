#!/usr/bin/env python
"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel table.

    python profiles/summarize.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.md

`steps` = number of bench steps the capture covered (for the per-step column).
ncu serialises launches and runs them cold-cache, so the SHARE column is what to compare
with the CUDA-event stage times printed by bench.py, not the absolute microseconds.
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    agg = collections.OrderedDict()
    for r in rows:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"]
        for cut in ("(", "<"):
            if name.startswith("void "):
                name = name[5:]
        name = name.split("(")[0]
        a = agg.setdefault(name, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", "")) / 1e3
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | us/launch | share | grid (first) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |" % (k, n, t, t / n, 100 * t / total, g, b))
    print("\ntotal %.1f us over %d launches" % (total, sum(a[0] for a in agg.values())) +
          (" = %.1f us/step over %g steps" % (total / steps, steps) if steps else ""))


if __name__ == "__main__":
    main()
