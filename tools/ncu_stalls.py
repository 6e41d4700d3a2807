#!/usr/bin/env python
"""Reads a `ncu --set full --import-source on` report here (no GPU needed) and prints, per profiled launch, the warp-stall
breakdown from the PC samples and the instructions that collected the most samples.

    python tools/ncu_stalls.py gpurun_out/tail_band_full.ncu-rep [top_n]

Two findings of round 1 came out of this view: the fully unrolled tail kernel was instruction-fetch bound (35 us -> 22 us
once its class loop was rolled), and the skeleton of `stem_tc_kernel` spends half of its samples in two
mbarrier try-wait loops (the producer / MMA hand-off chain)."""
import csv
import io
import subprocess
import sys


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, data = rows[0], rows[2:]
    pre = "smsp__pcsamp_warps_issue_stalled_"
    for r in data:
        d = dict(zip(hdr, r))
        stalls = {k[len(pre):]: float(v) for k, v in d.items() if k.startswith(pre) and not k.endswith("_not_issued") and v not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        print("%s  %s us  %s warp instructions, issue slots %s %% busy" % (
            d["Kernel Name"][:70], d.get("gpu__time_duration.sum", "?"), d.get("smsp__inst_executed.sum", "?"),
            d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "?")))
        print("   " + "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
    # source page: one table per profiled launch; only the first is summarised
    rows = list(csv.reader(io.StringIO(page(rep, "source"))))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not heads:
        return
    h0 = heads[0]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    hdr = rows[h0]
    body = [r for r in rows[h0 + 1:end] if len(r) == len(hdr)]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[si] or 0) for r in body) or 1
    print("first launch: %d SASS instructions (%d KB), %d samples; top %d:" % (len(body), len(body) * 16 // 1024, tot, top_n))
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top_n]:
        print("  %5.1f%%  executed %9s  %s" % (100.0 * int(r[si] or 0) / tot, r[ie], r[src].strip()[:100]))


if __name__ == "__main__":
    main()
