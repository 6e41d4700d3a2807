#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of a launch window.

    python tools/summarize_launches.py gpurun_out/launches_dff.csv [first] [last] [--top N]
"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    rows = load(args[0])
    first = int(args[1]) if len(args) > 1 else 0
    last = int(args[2]) if len(args) > 2 else len(rows)
    win = rows[first:last]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in win:
        name = r["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void accel::", "")
        t = float(r["Metric Value"]) / 1e3
        agg[name][0] += 1
        agg[name][1] += t
        tot += t
    print("launches %d..%d of %d: %d launches, %.1f us (cold-cache, serialised: compare shares)" % (first, last, len(rows), len(win), tot))
    print("%-44s %6s %11s %7s" % ("kernel", "count", "total us", "share"))
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %11.1f %6.1f%%" % (name, c, t, 100 * t / tot))
    if top_n:
        print("\nlongest %d launches (index in window, kernel, grid, us)" % top_n)
        for i, r in sorted(enumerate(win), key=lambda ir: -float(ir[1]["Metric Value"]))[:top_n]:
            print("%5d %-40s %-14s %9.1f" % (i, r["Kernel Name"].split("(")[0].replace("unnamed>::", "")[-40:], r["Grid Size"],
                                            float(r["Metric Value"]) / 1e3))


if __name__ == "__main__":
    main()
