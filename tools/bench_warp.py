#!/usr/bin/env python
"""Times accel_warp (CUDA events, rotating buffers > L2) for several flow fields, next to a plain copy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import engine as E  # noqa: E402


def timeit(fn, n=30):
    for k in range(6):
        fn(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(n):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = "cuda:0"
    C, h, w = 2048, 64, 128
    g = torch.Generator().manual_seed(0)
    src = [torch.randn(1, C, h, w, generator=g).to(dev) for _ in range(3)]
    dst = [torch.empty(1, C, h, w, device=dev) for _ in range(3)]
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    flows = {
        "zero": torch.zeros(1, 2, h, w),
        "const(3.3,-1.7)": torch.stack([torch.full((h, w), 3.3), torch.full((h, w), -1.7)])[None],
        "smooth": torch.stack([3.0 * torch.sin(yy / 9.0) + 0.01 * xx, 2.0 * torch.cos(xx / 13.0)])[None],
        "smooth+noise": torch.stack([3.0 * torch.sin(yy / 9.0) + 0.01 * xx, 2.0 * torch.cos(xx / 13.0)])[None]
        + torch.randn(1, 2, h, w, generator=g) * 0.3,
        "random(2px)": torch.randn(1, 2, h, w, generator=g) * 2.0,
    }
    nbytes = 2 * C * h * w * 4 + 2 * h * w * 4
    us = timeit(lambda k: dst[k % 3].copy_(src[k % 3]))
    print("env:", {k: v for k, v in os.environ.items() if k.startswith("ACCEL_WARP")})
    print("torch copy_            %7.2f us  %7.1f GB/s" % (us, 2 * C * h * w * 4 / us / 1e3))
    for name, fl in flows.items():
        fl = fl.contiguous().to(dev)
        us = timeit(lambda k: E.warp(src[k % 3], fl, dst[k % 3]))
        print("warp %-16s %7.2f us  %7.1f GB/s" % (name, us, nbytes / us / 1e3))
    # the fused kernel (accel_warp_split: fp32 NCHW + split-fp16 NHWC head operand in one pass): 3 x 64 MiB algorithmic
    hi = [torch.empty(h, w, C, dtype=torch.float16, device=dev) for _ in range(3)]
    lo = [torch.empty(h, w, C, dtype=torch.float16, device=dev) for _ in range(3)]
    fbytes = 3 * C * h * w * 4 + 2 * h * w * 4
    variants = [v for v in os.environ.get("BENCH_WARP_VARIANTS", "2:16").split(",") if v]
    for v in variants:
        per_sm, nst = v.split(":")[:2]
        os.environ["ACCEL_WARP_FUSED_PER_SM"], os.environ["ACCEL_WARP_FUSED_NST"] = per_sm, nst
        os.environ["ACCEL_WARP_FUSED_DEBUG"] = dbg = (v.split(":") + ["0", "1"])[2]
        os.environ["ACCEL_WARP_FUSED_TMAP"] = tmap = (v.split(":") + ["0", "1"])[3]
        for name in ("smooth", "smooth+noise", "random(2px)"):
            fl = flows[name].contiguous().to(dev)
            us = timeit(lambda k: E.warp_split(src[k % 3], fl, dst[k % 3], hi[k % 3], lo[k % 3]))
            print("warp_split per_sm=%s nst=%-2s dbg=%s tmap=%s %-14s %7.2f us  %7.1f GB/s" % (per_sm, nst, dbg, tmap, name, us, fbytes / us / 1e3))


if __name__ == "__main__":
    main()
