"""Reader / writer of MXNet `.params` checkpoints and the reference's loading helpers.

Mirrors lib/utils/load_model.py (load_checkpoint :4-30, load_checkpoint_multi :32-58, load_param :73-93,
load_param_multi :95-116) and lib/utils/save_model.py (save_checkpoint :4-20) -- same names, arguments and
`arg:` / `aux:` key handling -- with numpy arrays in place of mx.nd.NDArray.  What the reference delegates to
`mx.nd.load` / `mx.nd.save` is restated here from MXNet's published NDArray-list container [MXNet-ext]:

    uint64  0x112 (kMXAPINDArrayListMagic)      uint64 reserved (0)
    uint64  number of arrays, then each array:
        V2  uint32 0xF993FAC9, int32 storage type (0 = dense), uint32 ndim, int64 dims[ndim]
        V1  uint32 0xF993FAC8,                                  uint32 ndim, int64 dims[ndim]
        V0  (MXNet <= 0.10, the release the reference pins, README.md:30)  uint32 ndim, uint32 dims[ndim]
        then  int32 dev_type, int32 dev_id, int32 type_flag, raw little-endian data   (nothing if ndim == 0)
    uint64  number of names, then each: uint64 length, bytes

No `.params` file ships with the reference (`*.params` is git-ignored there), so this row is pinned only by
byte fixtures assembled by hand in tests/test_params_io.py and by write -> read round trips.
"""
from __future__ import annotations

import struct

import numpy as np

LIST_MAGIC = 0x112
V1_MAGIC = 0xF993FAC8
V2_MAGIC = 0xF993FAC9
# mshadow type flags
_DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_FLAGS = {np.dtype(v): k for k, v in _DTYPES.items()}


class ParamsFormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.buf, self.pos = memoryview(buf), 0

    def take(self, n):
        if self.pos + n > len(self.buf):
            raise ParamsFormatError("truncated .params file (wanted %d bytes at offset %d of %d)" % (n, self.pos, len(self.buf)))
        out = self.buf[self.pos:self.pos + n]
        self.pos += n
        return out

    def unpack(self, fmt):
        return struct.unpack("<" + fmt, self.take(struct.calcsize("<" + fmt)))


def _read_ndarray(r):
    (magic,) = r.unpack("I")
    if magic == V2_MAGIC:
        (stype,) = r.unpack("i")
        if stype != 0:
            raise ParamsFormatError("sparse NDArray (storage type %d) is not supported" % stype)
        (ndim,) = r.unpack("I")
        shape = r.unpack("%dq" % ndim) if ndim else ()
    elif magic == V1_MAGIC:
        (ndim,) = r.unpack("I")
        shape = r.unpack("%dq" % ndim) if ndim else ()
    else:                                   # legacy: the word just read is ndim, dims are uint32
        ndim = magic
        if ndim > 32:
            raise ParamsFormatError("implausible NDArray header 0x%08x" % magic)
        shape = r.unpack("%dI" % ndim) if ndim else ()
    if ndim == 0:
        return None                         # is_none(): nothing else was written
    _dev_type, _dev_id, type_flag = r.unpack("iii")
    if type_flag not in _DTYPES:
        raise ParamsFormatError("unknown type flag %d" % type_flag)
    dt = np.dtype(_DTYPES[type_flag]).newbyteorder("<")
    count = int(np.prod(shape, dtype=np.int64))
    data = np.frombuffer(r.take(count * dt.itemsize), dtype=dt, count=count)
    return data.reshape(shape).astype(dt.newbyteorder("="), copy=True)


def nd_load(fname):
    """mx.nd.load: dict name -> array when the file carries names, else a list."""
    with open(fname, "rb") as f:
        r = _Reader(f.read())
    header, _reserved = r.unpack("QQ")
    if header != LIST_MAGIC:
        raise ParamsFormatError("%s is not an MXNet NDArray list (magic 0x%x)" % (fname, header))
    (n,) = r.unpack("Q")
    arrays = [_read_ndarray(r) for _ in range(n)]
    (m,) = r.unpack("Q")
    names = []
    for _ in range(m):
        (ln,) = r.unpack("Q")
        names.append(bytes(r.take(ln)).decode("utf-8"))
    if m == 0:
        return arrays
    if m != n:
        raise ParamsFormatError("%d names for %d arrays" % (m, n))
    return dict(zip(names, arrays))


def nd_save(fname, data, version=2):
    """mx.nd.save for a dict (or list) of numpy arrays.  version 2 / 1 / 0 selects the per-array header."""
    if isinstance(data, dict):
        names, arrays = list(data.keys()), list(data.values())
    else:
        names, arrays = [], list(data)
    out = [struct.pack("<QQQ", LIST_MAGIC, 0, len(arrays))]
    for a in arrays:
        a = np.ascontiguousarray(a)
        if a.dtype not in _FLAGS:
            raise TypeError("dtype %s cannot be stored in a .params file" % a.dtype)
        if a.ndim == 0:
            a = a.reshape(1)
        if version == 2:
            out.append(struct.pack("<Ii", V2_MAGIC, 0))
            out.append(struct.pack("<I%dq" % a.ndim, a.ndim, *a.shape))
        elif version == 1:
            out.append(struct.pack("<I", V1_MAGIC))
            out.append(struct.pack("<I%dq" % a.ndim, a.ndim, *a.shape))
        elif version == 0:
            out.append(struct.pack("<I%dI" % a.ndim, a.ndim, *a.shape))
        else:
            raise ValueError("version must be 0, 1 or 2")
        out.append(struct.pack("<iii", 1, 0, _FLAGS[a.dtype]))              # cpu(0)
        out.append(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
    out.append(struct.pack("<Q", len(names)))
    for nm in names:
        b = nm.encode("utf-8")
        out.append(struct.pack("<Q", len(b)))
        out.append(b)
    with open(fname, "wb") as f:
        f.write(b"".join(out))


# ---- lib/utils/load_model.py -----------------------------------------------------------------------------
def load_checkpoint(prefix, epoch, argprefix=""):
    """(arg_params, aux_params) of '%s-%04d.params' % (prefix, epoch); keys lose their `arg:` / `aux:` tag
    and gain `argprefix` unless they already start with it (load_model.py:15-30)."""
    save_dict = nd_load("%s-%04d.params" % (prefix, epoch))
    arg_params, aux_params = {}, {}
    for k, v in save_dict.items():
        tp, name = k.split(":", 1)
        if name[:len(argprefix)] != argprefix:
            name = argprefix + name
        if tp == "arg":
            arg_params[name] = v
        if tp == "aux":
            aux_params[name] = v
    return arg_params, aux_params


def load_checkpoint_multi(prefix1, prefix2, epoch):
    """Two checkpoints merged, the second overriding the first (load_model.py:32-58)."""
    arg_params, aux_params = {}, {}
    for prefix in (prefix1, prefix2):
        for k, v in nd_load("%s-%04d.params" % (prefix, epoch)).items():
            tp, name = k.split(":", 1)
            if tp == "arg":
                arg_params[name] = v
            if tp == "aux":
                aux_params[name] = v
    return arg_params, aux_params


def _drop_test_suffix(arg_params):
    for test in [k for k in arg_params.keys() if "_test" in k]:               # load_model.py:89-92
        arg_params[test.replace("_test", "")] = arg_params.pop(test)


def load_param(prefix, epoch, convert=False, ctx=None, process=False, argprefix=""):
    """load_model.py:73-93.  `convert` / `ctx` are accepted for signature parity; arrays stay on the host
    (Engine.set_params uploads them)."""
    arg_params, aux_params = load_checkpoint(prefix, epoch, argprefix)
    if process:
        _drop_test_suffix(arg_params)
    return arg_params, aux_params


def load_param_multi(prefix1, prefix2, epoch, convert=False, ctx=None, process=False):
    arg_params, aux_params = load_checkpoint_multi(prefix1, prefix2, epoch)
    if process:
        _drop_test_suffix(arg_params)
    return arg_params, aux_params


def load_demo_params(model1_prefix, model2_prefix, epoch=0):
    """The four lines of dff_deeplab/demo.py:192-195: the Accel checkpoint, then the DeepLab-DCN one on top."""
    arg_params, aux_params = load_param(model1_prefix, epoch, process=True)
    arg_params_dcn, aux_params_dcn = load_param(model2_prefix, epoch, process=True)
    arg_params.update(arg_params_dcn)
    aux_params.update(aux_params_dcn)
    return arg_params, aux_params


# ---- lib/utils/save_model.py -----------------------------------------------------------------------------
def save_checkpoint(prefix, epoch, arg_params, aux_params, version=2):
    save_dict = {("arg:%s" % k): _to_numpy(v) for k, v in arg_params.items()}
    save_dict.update({("aux:%s" % k): _to_numpy(v) for k, v in aux_params.items()})
    nd_save("%s-%04d.params" % (prefix, epoch), save_dict, version=version)


def _to_numpy(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def split_arg_aux(params):
    """Synthetic parameter dicts are flat; BatchNorm moving statistics are `aux:` in MXNet."""
    arg, aux = {}, {}
    for k, v in params.items():
        (aux if k.endswith(("_moving_mean", "_moving_var")) else arg)[k] = v
    return arg, aux
