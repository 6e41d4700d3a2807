#!/usr/bin/env python
"""Is the fp32 accumulation inside tcgen05.mma biased (round-toward-zero) and how large is the conv error vs fp64?

One 3x3 conv layer (K = 9*Cin) through accel_conv_layer on the tcgen05 engine and on the CUDA-core FFMA engine, against
an fp64 reference computed with torch on the same GPU (test tooling).  All-positive operands make a truncation bias
visible as a negative mean signed error; zero-mean operands show the random part."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import engine as E  # noqa: E402


def probe(cin, cout, h, w, positive, k=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, cin, h, w, generator=g) if positive else torch.randn(1, cin, h, w, generator=g)
    wt = torch.rand(cout, cin, k, k, generator=g) if positive else torch.randn(cout, cin, k, k, generator=g)
    wt = wt / (cin * k * k) ** 0.5
    x, wt = x.cuda(), wt.cuda()
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), padding=k // 2)
    f32 = torch.nn.functional.conv2d(x, wt, padding=k // 2).double()
    out = {}
    for name, eng in (("tcgen05", 2), ("ffma", 1)):
        y = E.conv_layer(x, wt, "conv", 1, k // 2, 1, engine=eng).double()
        out[name] = y
    out["torch_fp32(cudnn)"] = f32
    scale = ref.abs().mean().item()
    print("cin %4d cout %4d %dx%d k%d %s: |ref| mean %.4g" % (cin, cout, h, w, k, "positive" if positive else "zero-mean", scale))
    for name, y in out.items():
        d = (y - ref)
        print("   %-18s mean signed err / |ref| = %+.3e   rms = %.3e   max = %.3e" % (
            name, (d.mean() / scale).item(), (d.pow(2).mean().sqrt() / scale).item(), (d.abs().max() / scale).item()))


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for positive in (True, False):
        probe(512, 256, 32, 64, positive)
        probe(64, 64, 32, 64, positive)
        probe(2048, 256, 32, 64, positive, k=1)
    print("---- single-group / short-K probes (1x1 conv)")
    for positive in (True, False):
        probe(16, 64, 32, 64, positive, k=1)
        probe(64, 64, 32, 64, positive, k=1)
        probe(256, 64, 32, 64, positive, k=1)
