#!/usr/bin/env python
"""Per-kernel count of the SASS opcodes that prove the Blackwell paths (no GPU needed):

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt

UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor loads / stores,
UBLKCP = cp.async.bulk (1-D TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier operations."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "accel_b200", "libaccel_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in ("UTCHMMA",):
            counts[kern]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif base in OPS:
            counts[kern][base] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# %s -- cuobjdump -sass opcode counts per kernel (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-64s " % "kernel" + " ".join("%12s" % o for o in OPS))
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), dem):
        short = re.sub(r"\(anonymous namespace\)::|accel::|void ", "", name).split("(")[0]
        print("%-64s " % short[:64] + " ".join("%12d" % c[o] for o in OPS))
        tot.update(c)
    print("%-64s " % "TOTAL" + " ".join("%12d" % tot[o] for o in OPS))


if __name__ == "__main__":
    main()
