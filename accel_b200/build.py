"""Builds accel_b200/libaccel_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension:
the library is a plain C ABI, see include/accel_b200.h)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaccel_b200.so")
SOURCES = ["abi.cu", "graph.cu", "nets.cu", "conv_ffma.cu", "conv_tc.cu", "stem_tc.cu", "kernels_basic.cu", "kernels_io.cu", "warp_staged.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False, trace=False):
    """trace=True: a second library, libaccel_b200_trace.so, with conv_tc_kernel's in-kernel timeline compiled in
    (tools/tc_trace.py selects it through ACCEL_B200_LIB)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "accel_b200.h"))
    objdir = os.path.join(HERE, "build_trace" if trace else "build")
    lib = LIB.replace(".so", "_trace.so") if trace else LIB
    flags = NVCC_FLAGS + (["-DACCEL_TC_TRACE_BUILD"] if trace else [])
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== nvcc %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(lib):
        cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", lib] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, trace="--trace" in sys.argv))
