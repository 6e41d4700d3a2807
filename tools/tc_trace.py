#!/usr/bin/env python
"""Runs one layer of tools/bench_layer.py with the in-kernel timeline on (ACCEL_TC_DEBUG bit 2048 + ACCEL_TC_TRACE) and
prints, per tile of CTA 0, where the cycles went: producer first/last TMA issue, MMA warp (accumulators free -> first stage
landed -> last MMA issued), epilogue warp (waiting -> accumulator complete -> released)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    layer = sys.argv[1] if len(sys.argv) > 1 else "res5_2a_x5"
    extra = dict(kv.split("=") for kv in sys.argv[2:])
    env = dict(os.environ, ACCEL_TC_TRACE="1", **extra)
    env.setdefault("ACCEL_B200_LIB", os.path.join(ROOT, "accel_b200", "libaccel_b200_trace.so"))   # python accel_b200/build.py --trace
    env["ACCEL_TC_DEBUG"] = str(int(extra.get("ACCEL_TC_DEBUG", "0")) | 2048)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_layer.py"), "--sweep", "env", "--set", layer],
                         env=env, capture_output=True, text=True).stderr
    ev = {}
    end = None
    for m in re.finditer(r"TC_TRACE tag (\d+) item (\d+) t (-?\d+)", out):
        tag, item, t = int(m.group(1)), int(m.group(2)), int(m.group(3))
        if tag == 9:
            end = t
        elif tag:
            ev.setdefault(item, {})[tag] = t
    print(layer, extra, [l for l in out.splitlines() if "ACCEL_LAYER" in l][-1:])
    print("%6s %8s %8s | %8s %8s %8s | %8s %8s %8s | mainloop  epi_wait->done  tile_period" % ("item", "tma0", "tmaN", "acc_free", "stage0", "mmaN", "epi_rdy", "acc_full", "released"))
    prev = None
    for item in sorted(ev):
        e = ev[item]
        g = lambda k: e.get(k, -1)
        period = g(3) - prev if prev is not None else 0
        prev = g(3)
        print("%6d %8d %8d | %8d %8d %8d | %8d %8d %8d | %8d %8d %8d" % (item, g(1), g(2), g(3), g(4), g(5), g(6), g(7), g(8), g(5) - g(3), g(8) - g(7), period))
    print("end", end)


if __name__ == "__main__":
    main()
