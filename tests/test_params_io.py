"""MXNet `.params` container (accel_b200/params_io.py): byte fixtures assembled here by hand from the published
layout (independent of the writer), write -> read round trips for the three per-array header versions, and the
reference's load_checkpoint / load_param key handling (lib/utils/load_model.py)."""
import struct

import numpy as np
import pytest

from accel_b200 import params_io as pio


def _hand_file(tmp_path, version):
    """Two named arrays laid out byte by byte: conv weight (2,1,1,3) float32 and a moving_var (2,) float32."""
    w = np.arange(6, dtype="<f4").reshape(2, 1, 1, 3) * 0.5 - 1.0
    v = np.array([1.5, 0.25], dtype="<f4")
    b = struct.pack("<QQQ", 0x112, 0, 2)
    for a in (w, v):
        if version == 2:
            b += struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape)
        elif version == 1:
            b += struct.pack("<I", 0xF993FAC8) + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape)
        else:
            b += struct.pack("<I", a.ndim) + struct.pack("<%dI" % a.ndim, *a.shape)
        b += struct.pack("<iii", 1, 0, 0) + a.tobytes()
    names = [b"arg:conv_test_weight", b"aux:bn_moving_var"]
    b += struct.pack("<Q", 2)
    for n in names:
        b += struct.pack("<Q", len(n)) + n
    path = tmp_path / ("hand_v%d-0007.params" % version)
    path.write_bytes(b)
    return str(tmp_path / ("hand_v%d" % version)), w, v


@pytest.mark.parametrize("version", [0, 1, 2])
def test_hand_assembled_file(tmp_path, version):
    prefix, w, v = _hand_file(tmp_path, version)
    d = pio.nd_load(prefix + "-0007.params")
    assert set(d) == {"arg:conv_test_weight", "aux:bn_moving_var"}
    assert d["arg:conv_test_weight"].shape == (2, 1, 1, 3) and np.array_equal(d["arg:conv_test_weight"], w)
    arg, aux = pio.load_checkpoint(prefix, 7)
    assert list(arg) == ["conv_test_weight"] and list(aux) == ["bn_moving_var"]
    assert np.array_equal(aux["bn_moving_var"], v)
    arg, aux = pio.load_param(prefix, 7, process=True)             # `_test` dropped (load_model.py:89-92)
    assert list(arg) == ["conv_weight"] and np.array_equal(arg["conv_weight"], w)
    arg, _ = pio.load_checkpoint(prefix, 7, argprefix="18_")       # prefix added unless already present
    assert list(arg) == ["18_conv_test_weight"]


@pytest.mark.parametrize("version", [0, 1, 2])
def test_round_trip_all_dtypes(tmp_path, version):
    rng = np.random.RandomState(version)
    data = {"arg:a": rng.randn(3, 4, 5).astype(np.float32), "arg:b": rng.randn(7).astype(np.float64),
            "aux:c": rng.randint(0, 255, (2, 2)).astype(np.uint8), "arg:d": rng.randint(-5, 5, (4,)).astype(np.int32),
            "arg:e": rng.randn(2, 3).astype(np.float16), "aux:f": np.arange(5, dtype=np.int64)}
    path = str(tmp_path / "rt.params")
    pio.nd_save(path, data, version=version)
    back = pio.nd_load(path)
    assert list(back) == list(data)
    for k in data:
        assert back[k].dtype == data[k].dtype and np.array_equal(back[k], data[k])
    pio.nd_save(path, list(data.values()), version=version)        # unnamed list form
    lst = pio.nd_load(path)
    assert isinstance(lst, list) and all(np.array_equal(x, y) for x, y in zip(lst, data.values()))


def test_checkpoint_round_trip_and_multi(tmp_path):
    from accel_b200 import synthetic
    params = {k: v.numpy() for k, v in synthetic.make_params("18").items()}
    arg, aux = pio.split_arg_aux(params)
    assert all(k.endswith(("_moving_mean", "_moving_var")) for k in aux) and len(aux) > 10
    n = len(arg) // 2
    keys = sorted(arg)
    pio.save_checkpoint(str(tmp_path / "m1"), 0, {k: arg[k] for k in keys[:n]}, aux)
    pio.save_checkpoint(str(tmp_path / "m2"), 0, {k: arg[k] for k in keys[n - 3:]}, {}, version=0)
    a, x = pio.load_param_multi(str(tmp_path / "m1"), str(tmp_path / "m2"), 0)
    assert set(a) == set(arg) and set(x) == set(aux)
    assert all(np.array_equal(a[k], arg[k]) for k in arg) and all(np.array_equal(x[k], aux[k]) for k in aux)
    a2, x2 = pio.load_demo_params(str(tmp_path / "m1"), str(tmp_path / "m2"))
    assert set(a2) == set(arg) and set(x2) == set(aux)


def test_errors(tmp_path):
    p = tmp_path / "bad.params"
    p.write_bytes(struct.pack("<QQQ", 0x113, 0, 0))
    with pytest.raises(pio.ParamsFormatError):
        pio.nd_load(str(p))
    p.write_bytes(struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<Iq", 1, 100) +
                  struct.pack("<iii", 1, 0, 0) + b"\x00" * 8)
    with pytest.raises(pio.ParamsFormatError):
        pio.nd_load(str(p))                                         # truncated payload
    with pytest.raises(TypeError):
        pio.nd_save(str(p), {"a": np.zeros(2, dtype=np.complex64)})
