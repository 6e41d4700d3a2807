#!/usr/bin/env python
"""Summarises a saved `ncu --set full --import-source on` report of conv_tc_kernel per warp role: stall reasons of the
whole kernel, dynamic warp instructions and stall samples of the producer / MMA / epilogue code regions (found by their
landmark opcodes), and the hottest instructions.  Usage: python tools/ncu_regions.py report.ncu-rep [tiles]"""
import csv
import subprocess
import sys
from collections import Counter

NCU = "/usr/local/cuda/bin/ncu"


def page(rep, name, extra=()):
    out = subprocess.run([NCU, "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    d = dict(zip(raw[0], raw[2]))
    print("kernel time %s us, regs %s, tensor pipe %s %%, issue active %s %%, dram %s %%" % (
        d.get("gpu__time_duration.sum"), d.get("launch__registers_per_thread"),
        d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    st = [(int(v), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, v in d.items()
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
    tot = sum(n for n, _ in st)
    print("stall reasons:", ", ".join("%s %.1f%%" % (h, 100.0 * n / tot) for n, h in sorted(st, reverse=True)[:8]))
    src = page(rep, "source", ("--print-source", "sass"))
    for i, r in enumerate(src):
        if "Source" in r and any("Sampl" in c for c in r):
            hdr, start = r, i + 1
            break
    blk = [r for r in src[start:] if len(r) > 5]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    n = lambda r, j: int(r[j]) if r[j].isdigit() else 0
    print("instructions %d, executed warp-instr %d, samples %d" % (len(blk), sum(n(r, ie) for r in blk), sum(n(r, isamp) for r in blk)))
    c = Counter()
    for r in blk:
        s = r[1].strip()
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        c[op] += n(r, ie)
    print("executed by opcode:", ", ".join("%s %d" % kv for kv in c.most_common(14)))
    print("hottest instructions (samples, executed, sass):")
    for k, r in sorted(enumerate(blk), key=lambda kr: -n(kr[1], isamp))[:14]:
        print("  %5d %6d %9d  %s" % (k, n(r, isamp), n(r, ie), r[1].strip()[:90]))


if __name__ == "__main__":
    main()
