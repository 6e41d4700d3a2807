"""C-ABI library: loads, exports every declared symbol, enumerates parameters without a GPU, and
refuses to compute without one (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from accel_b200 import _lib, netspec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "accel_b200.h")).read()
    declared = set(re.findall(r"\b(accel_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def _create(lib, version, h=128, w=256, flags=0):
    cfg = _lib.AccelConfig(_lib.VERSION_CODE[version], h, w, 19, 0, flags)
    handle = C.c_void_p()
    rc = lib.accel_create(C.byref(cfg), C.byref(handle))
    return rc, handle


@pytest.mark.parametrize("version", netspec.VERSIONS)
def test_param_inventory_matches_python_spec(lib, version):
    rc, h = _create(lib, version)
    assert rc == 0, lib.accel_last_error(None)
    got = {}
    name, shape, ndim = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
    for i in range(lib.accel_param_count(h)):
        assert lib.accel_param_info(h, i, C.byref(name), shape, C.byref(ndim)) == 0
        got[name.value.decode()] = tuple(shape[j] for j in range(ndim.value))
    want = {k: tuple(v[0]) for k, v in netspec.param_spec(version).items()}
    assert set(got) == set(want), (sorted(set(got) - set(want))[:5], sorted(set(want) - set(got))[:5])
    for k in want:
        assert got[k] == want[k], k
    lib.accel_destroy(h)


def test_create_rejects_bad_config(lib):
    rc, _ = _create(lib, "18", h=100, w=256)
    assert rc != 0 and b"multiples of 128" in lib.accel_last_error(None)
    cfg = _lib.AccelConfig(77, 128, 256, 19, 0, 0)
    handle = C.c_void_p()
    assert lib.accel_create(C.byref(cfg), C.byref(handle)) != 0


def test_set_param_checks_names_and_shapes(lib):
    rc, h = _create(lib, "dff")
    assert rc == 0
    import numpy as np
    a = np.zeros((1024, 2048, 1, 1), dtype=np.float32)
    shp = (C.c_int64 * 4)(*a.shape)
    assert lib.accel_set_param(h, b"fc6_weight", a.ctypes.data_as(C.c_void_p), shp, 4) == 0
    assert lib.accel_set_param(h, b"no_such_weight", a.ctypes.data_as(C.c_void_p), shp, 4) != 0
    assert b"unknown parameter" in lib.accel_last_error(h)
    bad = (C.c_int64 * 4)(1024, 2047, 1, 1)
    assert lib.accel_set_param(h, b"fc6_weight", a.ctypes.data_as(C.c_void_p), bad, 4) != 0
    assert b"shape mismatch" in lib.accel_last_error(h)
    lib.accel_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    rc, h = _create(lib, "dff")
    assert rc == 0
    assert lib.accel_finalize(h) != 0
    assert b"no CUDA device" in lib.accel_last_error(h) or b"never set" in lib.accel_last_error(h)
    buf = (C.c_float * 16)()
    assert lib.accel_warp(buf, buf, (C.c_float * 16)(), 1, 2, 2, None) != 0
    err = C.create_string_buffer(256)
    assert lib.accel_head(buf, 4, 2, 2, buf, buf, 2, buf, buf, 2, buf, 0, err, 256) == 6 and b"CUDA" in err.value
    assert lib.accel_conv_layer(0, buf, 1, 4, 4, buf, 1, 1, 1, 0, 1, 1, None, None, None, 0, None, 0, buf, 0, err, 256) == 6
    lib.accel_destroy(h)
    from accel_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine("dff", 128, 256)


@pytest.mark.parametrize("version", ["dff", "18", "101"])
def test_interval_plan_builds_on_cpu_and_adds_no_parameters(lib, version):
    """accel_plan_interval is pure host logic until finalize: it must build without a GPU, enumerate exactly the
    parameters of the frame-by-frame plans (same networks, same names), and reject bad intervals / double planning."""
    rc, h = _create(lib, version)
    assert rc == 0
    n0 = lib.accel_param_count(h)
    assert lib.accel_plan_interval(h, 1) != 0 and b"interval must be" in lib.accel_last_error(h)
    assert lib.accel_plan_interval(h, 17) != 0
    assert lib.accel_plan_interval(h, 5) == 0, lib.accel_last_error(h)
    assert lib.accel_param_count(h) == n0
    assert lib.accel_plan_interval(h, 5) != 0 and b"already planned" in lib.accel_last_error(h)
    lib.accel_destroy(h)


def test_interval_forward_needs_a_plan(lib):
    rc, h = _create(lib, "dff")
    assert rc == 0
    ptrs = (C.c_void_p * 2)()
    assert lib.accel_interval_forward(h, ptrs, None, ptrs, None) != 0
    assert b"accel_plan_interval" in lib.accel_last_error(h)
    assert lib.accel_rbranch_forward(h, C.c_void_p(16), None, None, None) != 0
    assert b"correction network" in lib.accel_last_error(h)
    lib.accel_destroy(h)


def test_every_environment_switch_is_documented():
    """Every ACCEL_* variable the library or the host package reads appears in DESIGN.md (section 5.6), README.md or
    INTEGRATION.md: the switches are the record of what was measured and not adopted."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "accel_b200", "csrc", "*")) + glob.glob(os.path.join(root, "accel_b200", "*.py")):
        names |= set(re.findall(r'"(ACCEL_[A-Z0-9_]+)"', open(f).read()))
    docs = "".join(open(os.path.join(root, d)).read() for d in ("DESIGN.md", "README.md", "INTEGRATION.md"))
    missing = sorted(n for n in names if n not in docs and not any(n.startswith(p[:-1]) and p.endswith("*") for p in re.findall(r"ACCEL_[A-Z0-9_]+\*", docs)))
    # `ACCEL_WARP_FUSED_PER_SM`, `_NST`, ... are listed with a shared prefix
    missing = [n for n in missing if not (n.startswith("ACCEL_WARP_FUSED_") and "ACCEL_WARP_FUSED_PER_SM" in docs and ("`_" + n[len("ACCEL_WARP_FUSED_"):]) in docs)]
    assert not missing, missing
