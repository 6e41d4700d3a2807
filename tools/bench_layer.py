#!/usr/bin/env python
"""Times single layers through accel_conv_layer under different planner knobs (tuning aid, GPU only).

    python tools/bench_layer.py [--set NAME] 2> layer_times.txt

Each line printed by the library (stderr, prefix ACCEL_LAYER) is the best-of-n device time of the
contraction kernels of that layer; the knobs in effect are echoed before it."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from accel_b200 import engine as E  # noqa: E402

# name: (cin, cout, h, w, k, stride, pad, dil, residual)
LAYERS = {
    "res2_2c": (64, 256, 256, 512, 1, 1, 0, 1, True),
    "res2_2b": (64, 64, 256, 512, 3, 1, 1, 1, False),
    "res3_2c": (128, 512, 128, 256, 1, 1, 0, 1, True),
    "res3_2b": (128, 128, 128, 256, 3, 1, 1, 1, False),
    "res4_2a": (1024, 256, 64, 128, 1, 1, 0, 1, False),
    "res4_2b": (256, 256, 64, 128, 3, 1, 1, 1, False),
    "res4_2c": (256, 1024, 64, 128, 1, 1, 0, 1, True),
    "res5_2a": (2048, 512, 64, 128, 1, 1, 0, 1, False),
    "res5_2c": (512, 2048, 64, 128, 1, 1, 0, 1, True),
    "fc6": (2048, 1024, 64, 128, 1, 1, 0, 1, False),
    "res5_off": (512, 18, 64, 128, 3, 1, 1, 1, False),
    "flow_conv3_1": (256, 256, 128, 256, 3, 1, 1, 1, False),
    "flow_conv3": (128, 256, 128, 256, 5, 2, 2, 1, False),
    "flow_conv2": (64, 128, 512, 1024, 5, 2, 2, 1, False),
    "flow_conv4": (256, 512, 128, 256, 3, 2, 1, 1, False),
    "flow_conv5": (512, 512, 64, 128, 3, 2, 1, 1, False),
    "flow_conv6": (512, 1024, 32, 64, 3, 2, 1, 1, False),
    "r18_s1": (64, 64, 256, 512, 3, 1, 1, 1, True),
    "r18_s1_x4": (64, 64, 1024, 512, 3, 1, 1, 1, True),
    "r18_s2_x4": (128, 128, 512, 256, 3, 1, 1, 1, True),
    "r18_s3_x4": (256, 256, 256, 128, 3, 1, 1, 1, True),
    "r18_s4_x4": (512, 512, 256, 128, 3, 1, 2, 2, True),
    # the same layers over 5 frames stacked along H: what batching an interval's frames through one launch would give
    "res2_2c_x5": (64, 256, 1280, 512, 1, 1, 0, 1, True),
    "res2_br1_x5": (64, 256, 1280, 512, 1, 1, 0, 1, False),
    "res3_2c_x5": (128, 512, 640, 256, 1, 1, 0, 1, True),
    "res2_2a_x5": (256, 64, 1280, 512, 1, 1, 0, 1, False),
    "res2_2a": (256, 64, 256, 512, 1, 1, 0, 1, False),
    "res3_2a_x5": (512, 128, 640, 256, 1, 1, 0, 1, False),
    "res3_2a": (512, 128, 128, 256, 1, 1, 0, 1, False),
    "res3_2b_x5": (128, 128, 640, 256, 3, 1, 1, 1, False),
    "res2_2b_x5": (64, 64, 1280, 512, 3, 1, 1, 1, False),
    "res4_2a_x5": (1024, 256, 320, 128, 1, 1, 0, 1, False),
    "res4_2b_x5": (256, 256, 320, 128, 3, 1, 1, 1, False),
    "res4_2c_x5": (256, 1024, 320, 128, 1, 1, 0, 1, True),
    "res5_br1_x5": (1024, 2048, 320, 128, 1, 1, 0, 1, False),
    "res5_2a_x5": (2048, 512, 320, 128, 1, 1, 0, 1, False),
    "res5_2c_x5": (512, 2048, 320, 128, 1, 1, 0, 1, True),
    "fc6_x5": (2048, 1024, 320, 128, 1, 1, 0, 1, False),
    "flow_conv4_1": (512, 512, 32, 64, 3, 1, 1, 1, False),
    "flow_conv4_1_x4": (512, 512, 128, 64, 3, 1, 1, 1, False),
    "flow_conv5_1": (512, 512, 16, 32, 3, 1, 1, 1, False),
    "flow_conv5_1_x4": (512, 512, 64, 32, 3, 1, 1, 1, False),
    "flow_conv6_1_x4": (1024, 1024, 32, 16, 3, 1, 1, 1, False),
    "flow_conv6_1": (1024, 1024, 8, 16, 3, 1, 1, 1, False),
}
SWEEP = [
    {},
    {"ACCEL_TC_KROT": "0"},
    {"ACCEL_TC_BN": "128"},
    {"ACCEL_TC_BN": "128", "ACCEL_TC_KROT": "0"},
    {"ACCEL_TC_BN": "64"},
    {"ACCEL_TC_BN": "256", "ACCEL_TC_SPLITS": "2"},
    {"ACCEL_TC_BN": "128", "ACCEL_TC_SPLITS": "2"},
]


PAIR_SWEEP = [
    {},
    {"ACCEL_TC_PAIR": "1"},
    {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "256"},
    {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "128"},
    {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "256", "ACCEL_TC_SPLITS": "2"},
    {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "256", "ACCEL_TC_SPLITS": "3"},
    {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "128", "ACCEL_TC_SPLITS": "2"},
]

ONE = [{}]
EPI_SWEEP = [{}, {"ACCEL_TC_TMA_OUT": "0"}, {"ACCEL_TC_BN": "256"}, {"ACCEL_TC_BN": "256", "ACCEL_TC_TMA_OUT": "0"}, {"ACCEL_TC_BN": "64"}]

R02_SWEEP = [
    {"ACCEL_TC_ASLAB": "0"},
    {},
    {"ACCEL_TC_ASLAB_SA": "3"},
    {"ACCEL_TC_ASLAB_SA": "1"},
    {"ACCEL_TC_ASLAB_BO": "1"},
]

DEBUG_SWEEP = [
    {},
    {"ACCEL_TC_TMA_OUT": "0"},
    {"ACCEL_TC_BN": "128", "ACCEL_TC_TMA_OUT": "0"},
    {"ACCEL_TC_BN": "64"},
    {"ACCEL_TC_BN": "64", "ACCEL_TC_TMA_OUT": "0"},
    {"ACCEL_TC_DEBUG": "1"},
    {"ACCEL_TC_DEBUG": "2"},
    {"ACCEL_TC_DEBUG": "3"},
    {"ACCEL_TC_DEBUG": "7"},
    {"ACCEL_TC_BN": "128"},
    {"ACCEL_TC_BN": "128", "ACCEL_TC_DEBUG": "3"},
    {"ACCEL_TC_BN": "128", "ACCEL_TC_DEBUG": "7"},
]

PF_SWEEP = [{"ACCEL_TC_RES_PREFETCH": "0"}, {"ACCEL_TC_RES_PREFETCH": "1"}]
MMA_SWEEP = [{"ACCEL_TC_DEBUG": "256"}, {"ACCEL_TC_DEBUG": "256", "ACCEL_TC_NCAT": "0"},
             {"ACCEL_TC_DEBUG": "256", "ACCEL_TC_BN": "256", "ACCEL_TC_WIDE_KMAX": "100000"},
             {"ACCEL_TC_DEBUG": "256", "ACCEL_TC_BN": "64"}, {"ACCEL_TC_DEBUG": "256", "ACCEL_TC_BN": "64", "ACCEL_TC_NCAT": "0"},
             {}, {"ACCEL_TC_NCAT": "0"}, {"ACCEL_TC_BN": "256", "ACCEL_TC_WIDE_KMAX": "100000"}, {"ACCEL_TC_BN": "64"}]
_PW = {"ACCEL_TC_PAIR": "1", "ACCEL_TC_BN": "256", "ACCEL_TC_WIDE_KMAX": "100000"}
_SW = {"ACCEL_TC_BN": "256", "ACCEL_TC_WIDE_KMAX": "100000", "ACCEL_TC_NCAT": "0"}
PAIR2_SWEEP = [{}, dict(_SW), dict(_SW, ACCEL_TC_DEBUG="256"), dict(_SW, ACCEL_TC_DEBUG="128"),
               dict(_PW), dict(_PW, ACCEL_TC_DEBUG="256"), dict(_PW, ACCEL_TC_DEBUG="128"), dict(_PW, ACCEL_TC_DEBUG="384")]
SHORTK_SWEEP = [{}, {"ACCEL_TC_DEBUG": "496"}, {"ACCEL_TC_DEBUG": "240"}, {"ACCEL_TC_DEBUG": "256"}, {"ACCEL_TC_BN": "256"}, {"ACCEL_TC_BN": "256", "ACCEL_TC_DEBUG": "496"},
                {"ACCEL_TC_BN": "64"}, {"ACCEL_TC_TMA_OUT": "0"}, {"ACCEL_TC_BN": "256", "ACCEL_TC_TMA_OUT": "0"}, {"ACCEL_TC_NCAT": "0"}]
LD_SWEEP = [{}] + [{"ACCEL_TC_DEBUG": str(b)} for b in (128, 240, 240 + 256, 240 + 512, 240 + 1024, 256, 512, 1024)]
EPI2_SWEEP = [{}] + [{"ACCEL_TC_DEBUG": str(b)} for b in (16, 32, 64, 128, 48, 112, 240, 128 + 16, 128 + 32, 128 + 64)]


S2_SWEEP = [{}, {"ACCEL_TC_CHAINS": "0"}, {"ACCEL_TC_TMA_OUT": "1"}, {"ACCEL_TC_TMA_OUT": "1", "ACCEL_TC_CHAINS": "0"},
            {"ACCEL_TC_ASLAB": "1"}, {"ACCEL_TC_ASLAB": "1", "ACCEL_TC_ASLAB_SA": "3"}, {"ACCEL_TC_BN": "64"},
            {"ACCEL_TC_DEBUG": "128"}, {"ACCEL_TC_DEBUG": "256"}, {"ACCEL_TC_DEBUG": "384"}, {"ACCEL_TC_DEBUG": "391"},
            {"ACCEL_TC_DEBUG": "391", "ACCEL_TC_CHAINS": "0"}, {"ACCEL_TC_DEBUG": "7"}, {"ACCEL_TC_DEBUG": "1"}]


PAIR3_SWEEP = [{"ACCEL_TC_PAIR": "0"}, {"ACCEL_TC_PAIR": "1"}, {"ACCEL_TC_PAIR": "1", "ACCEL_TC_CHAINS": "0"}, {"ACCEL_TC_PAIR": "0", "ACCEL_TC_CHAINS": "0"},
               {"ACCEL_TC_PAIR": "1", "ACCEL_TC_DEBUG": "128"}, {"ACCEL_TC_PAIR": "1", "ACCEL_TC_DEBUG": "256"}]
PAIR4_SWEEP = [{"ACCEL_TC_PAIR": "1", "ACCEL_TC_DEBUG": str(d)} for d in (128, 384, 128 + 512, 128 + 1024, 391)] + \
              [{"ACCEL_TC_PAIR": "0", "ACCEL_TC_DEBUG": str(d)} for d in (128, 384, 128 + 512, 128 + 1024)]
MC_SWEEP = [{"ACCEL_TC_MCAST": "0"}, {"ACCEL_TC_MCAST": "1"}, {"ACCEL_TC_MCAST": "1", "ACCEL_TC_DEBUG": "128"}, {"ACCEL_TC_MCAST": "0", "ACCEL_TC_DEBUG": "128"}]
SPIN_SWEEP = [{}, {"ACCEL_TC_DEBUG": "8192"}, {"ACCEL_TC_DEBUG": "384"}, {"ACCEL_TC_DEBUG": str(384 + 8192)}]
E1_SWEEP = [{"ACCEL_TC_EPI1": "0"}, {"ACCEL_TC_EPI1": "1"}, {"ACCEL_TC_EPI1": "0", "ACCEL_TC_EPIW16": "0"}]
T0_SWEEP = [{}, {"ACCEL_TC_TMA_OUT": "0"}, {"ACCEL_TC_EPIW16": "1"}, {"ACCEL_TC_DEBUG": "128"}, {"ACCEL_TC_DEBUG": "256"}]
W16_SWEEP = [{"ACCEL_TC_EPIW16": "0"}, {"ACCEL_TC_EPIW16": "1"}]
RA_SWEEP = [{}] + [{"ACCEL_TC_RES_AHEAD": str(d)} for d in (2, 4, 6, 8, 12)]
PF2_SWEEP = [{}, {"ACCEL_TC_PREFETCH": "4"}, {"ACCEL_TC_PREFETCH": "8"}, {"ACCEL_TC_PREFETCH": "260"}, {"ACCEL_TC_PREFETCH": "264"}, {"ACCEL_TC_PREFETCH": "2"}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sweep", default="default")
    a = ap.parse_args()
    os.environ["ACCEL_LAYER_REPS"] = str(a.reps)
    names = [n for n in LAYERS if not a.set or n in a.set.split(",")]
    dev = "cuda:0"
    for name in names:
        cin, cout, h, w, k, s, p, d, res = LAYERS[name]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(1, cin, h, w, generator=g).to(dev)
        wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        ho = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
        wo = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
        r = torch.randn(1, cout, ho, wo, generator=g).to(dev) if res else None
        if a.sweep == "env":                      # whatever the caller's environment says, untouched (tools/tc_trace.py)
            sys.stderr.write("%-14s env " % name)
            sys.stderr.flush()
            E.conv_layer(x, wt, "conv", s, p, d, act=1, residual=r, engine=2)
            continue
        for knobs in {"debug": DEBUG_SWEEP, "pair": PAIR_SWEEP, "r02": R02_SWEEP, "one": ONE, "epi": EPI_SWEEP, "pf": PF_SWEEP, "epi2": EPI2_SWEEP, "ld": LD_SWEEP, "mma": MMA_SWEEP, "pair2": PAIR2_SWEEP, "shortk": SHORTK_SWEEP, "s2": S2_SWEEP, "pf2": PF2_SWEEP, "ra": RA_SWEEP, "w16": W16_SWEEP, "t0": T0_SWEEP, "e1": E1_SWEEP, "spin": SPIN_SWEEP, "mc": MC_SWEEP, "pair4": PAIR4_SWEEP, "pair3": PAIR3_SWEEP, "env": None}.get(a.sweep, SWEEP):
            for kk in ("ACCEL_TC_BN", "ACCEL_TC_SPLITS", "ACCEL_TC_KROT", "ACCEL_TC_STAGES", "ACCEL_TC_DEBUG", "ACCEL_TC_TMA_OUT", "ACCEL_TC_PAIR", "ACCEL_TC_ASLAB",
                       "ACCEL_TC_EPI1", "ACCEL_TC_EPIW16", "ACCEL_TC_MCAST", "ACCEL_TC_PREFETCH", "ACCEL_TC_RES_AHEAD", "ACCEL_TC_CHAINS", "ACCEL_TC_ASLAB_SA", "ACCEL_TC_ASLAB_BO", "ACCEL_TC_RES_PREFETCH", "ACCEL_TC_NCAT", "ACCEL_TC_WIDE_KMAX"):
                os.environ.pop(kk, None)
            os.environ.update(knobs)
            sys.stderr.write("%-14s %-44s " % (name, knobs))
            sys.stderr.flush()
            try:
                E.conv_layer(x, wt, "conv", s, p, d, act=1, residual=r, engine=2)
            except RuntimeError as e:
                sys.stderr.write("FAILED %s\n" % e)


if __name__ == "__main__":
    main()
