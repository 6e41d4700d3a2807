"""Host-side mirror of the reference's inference interface for this path, so that the loop in
dff_deeplab/demo.py:228-250 (and pred_eval, core/tester.py:246-256) runs unchanged in shape:

    sym_instance = accel_18()                                     # demo.py:127-129
    key_sym = sym_instance.get_key_test_symbol(config)
    cur_sym = sym_instance.get_cur_test_symbol(config)
    key_predictor = Predictor(key_sym, data_names, label_names, context=[gpu(0)], ..., arg_params, aux_params)
    cur_predictor = Predictor(cur_sym, ...)
    output_all, feat = im_segment(key_predictor, data_batch)      # tester.py:158-171

Same names, same argument meaning, same output-name dictionaries (accel_18.py:157,237;
accel_101.py:191).  Tensors are CUDA torch tensors instead of mx.nd.NDArray; a Predictor owns its
output tensors and reuses them on every forward, as an MXNet executor does.  There is no CPU path.
"""
from __future__ import annotations

import torch

from .engine import Engine
from .netspec import FEAT_DIM, NUM_CLASSES


class gpu:                                            # mx.gpu(i)
    def __init__(self, device_id=0):
        self.device_id = int(device_id)


class Symbol:
    """What get_key_test_symbol / get_cur_test_symbol return: which graph, and its output names."""

    def __init__(self, version, kind):
        self.version, self.kind = str(version), kind

    def list_outputs(self):
        if self.kind == "deeplab":                    # deeplab/symbols/resnet_v1_101_deeplab.py:798-803
            return ["softmax_output"]
        if self.kind == "key":                        # accel_18.py:157
            return ["data_key", "feat_key", "res5c_relu_output", "croped_score_output"]
        score = "croped_score_output" if self.version in ("101", "dff") else "correction_output"
        return ["data_key", "warping_feat_output", score]          # accel_18.py:237, accel_101.py:191


class deeplab:
    """deeplab/symbols/resnet_v1_101_deeplab.py get_symbol(is_train=False) shaped stand-in for BASELINE config 1: the
    correction network of Accel-<version> alone (its own DeepLab head) on one frame.  Output names follow the test
    symbol (:798-803): `softmax_output`; deeplab/core/tester.py:84-85 takes its argmax."""

    def __init__(self, version="18"):
        self.version = str(version)

    def get_symbol(self, cfg=None, is_train=False):
        if is_train:
            raise NotImplementedError("training graphs are out of scope (SURVEY.md section 8)")
        return Symbol(self.version, "deeplab")


class _AccelSymbols:
    version = None

    def get_key_test_symbol(self, cfg=None):
        return Symbol(self.version, "key")

    def get_cur_test_symbol(self, cfg=None):
        return Symbol(self.version, "cur")


class accel_18(_AccelSymbols):
    version = "18"


class accel_34(_AccelSymbols):
    version = "34"


class accel_50(_AccelSymbols):
    version = "50"


class accel_101(_AccelSymbols):
    version = "101"


class dff_deeplab(_AccelSymbols):                     # FlowNet + warp + task head only (Deep Feature Flow)
    version = "dff"


class DataBatch:
    """Minimal stand-in for mx.io.DataBatch: data[0] = [data, data_key, feat_key] (demo.py:184,222-225)."""

    def __init__(self, data, label=None, pad=0, index=0, provide_data=None, provide_label=None):
        self.data, self.label, self.pad, self.index = data, label or [], pad, index
        self.provide_data, self.provide_label = provide_data, provide_label


_ENGINES = {}


def _fingerprint(*dicts):
    """Content fingerprint of the parameter dicts: names, shapes and a CRC-32 of every array's bytes.  Keying the
    engine cache on id() alone would silently reuse stale device weights after `arg_params.update(other_ckpt)`, or
    when a freed dict's id is recycled."""
    import zlib

    import numpy as np
    crc = 0
    n = 0
    for d in dicts:
        for name in sorted(d or {}):
            v = d[name]
            a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            a = np.ascontiguousarray(a)
            crc = zlib.crc32(name.encode(), crc)
            crc = zlib.crc32(repr((a.shape, str(a.dtype))).encode(), crc)
            crc = zlib.crc32(memoryview(a).cast("B"), crc)
            n += 1
    return (n, crc & 0xFFFFFFFF)


def _shared_engine(version, height, width, device, arg_params, aux_params, flags):
    # The key and cur Predictors of one demo share weights (same arg_params contents): one handle serves both.
    key = (version, height, width, device, _fingerprint(arg_params, aux_params), flags)
    eng = _ENGINES.get(key)
    if eng is None:
        params = dict(arg_params or {})
        params.update(aux_params or {})
        eng = Engine(version, height, width, params=params, device=device, flags=flags)
        _ENGINES.clear()                              # keep one live engine set; weights are large
        _ENGINES[key] = eng
    return eng


class Predictor:
    """dff_deeplab/core/tester.py:22-35."""

    def __init__(self, symbol, data_names, label_names, context=None, max_data_shapes=None, provide_data=None,
                 provide_label=None, arg_params=None, aux_params=None, engine=None, emit_scores=True, flags=0):
        names = list(data_names)
        if symbol.kind == "deeplab":                  # deeplab/function/test_deeplab.py:66-75: data_names = ['data']
            if "data" not in names:
                raise ValueError("data_names must contain 'data'")
            names = names + [n for n in ("data_key", "feat_key") if n not in names]
        if not all(k in names for k in ("data", "data_key", "feat_key")):
            raise ValueError("data_names must contain 'data', 'data_key', 'feat_key' (demo.py:184; TestLoader adds "
                             "'im_info', core/loader.py:214)")
        self._slots = tuple(names.index(k) for k in ("data", "data_key", "feat_key"))
        self.symbol = symbol
        self.output_names = symbol.list_outputs()
        ctx = context[0] if isinstance(context, (list, tuple)) else context
        device = getattr(ctx, "device_id", 0) if ctx is not None else 0
        shapes = dict(provide_data[0]) if provide_data else dict(max_data_shapes[0])
        _, _, h, w = shapes["data"]
        if engine is None:
            engine = _shared_engine(symbol.version, h, w, device, arg_params, aux_params, flags)
        self.engine = engine
        self.emit_scores = emit_scores
        dev = engine.torch_device
        self._feat = [torch.empty(engine.feat_shape, device=dev) for _ in range(2)]
        self._flip = 0
        self._score = torch.empty(1, NUM_CLASSES, h, w, device=dev) if emit_scores else None
        self._label = torch.empty(h, w, dtype=torch.uint8, device=dev)

    def predict(self, data_batch):
        """Returns [ {output_name: tensor} ] for the single device, like tester.py:32-35."""
        eng = self.engine
        if self.symbol.kind == "deeplab":
            data = data_batch.data[0][self._slots[0]]
            eng.rbranch_forward(data, self._score, self._label)
            out = {"label_output": self._label}
            if self.emit_scores:                      # SoftmaxOutput at inference = softmax over classes (:800)
                out["croped_score_output"] = self._score
                out["softmax_output"] = torch.softmax(self._score, dim=1)
            return [out]
        data, data_key, feat_key = (data_batch.data[0][i] for i in self._slots)
        feat_out = self._feat[self._flip]
        if self.symbol.kind == "cur" and feat_key.data_ptr() == feat_out.data_ptr():
            self._flip ^= 1                           # chained schedule feeds our own output back in
            feat_out = self._feat[self._flip]
        if self.symbol.kind == "key":
            eng.key_forward(data, feat_out, self._score, self._label)
            out = {"data_key": data_key, "feat_key": feat_key, "res5c_relu_output": feat_out}
        else:
            if tuple(feat_key.shape) != eng.feat_shape:
                raise ValueError("feat_key has shape %s, expected %s" % (tuple(feat_key.shape), eng.feat_shape))
            eng.cur_forward(data, data_key, feat_key, feat_out, self._score, self._label)
            out = {"data_key": data_key, "warping_feat_output": feat_out}
        if self.emit_scores:
            out[self.output_names[-1]] = self._score
        out["label_output"] = self._label             # extra: uint8 argmax, so callers can skip the volume
        return [out]


def im_segment(predictor, data_batch):
    """dff_deeplab/core/tester.py:158-171."""
    output_all = predictor.predict(data_batch)
    if "res5c_relu_output" in output_all[0]:
        feat = output_all[0]["res5c_relu_output"]
    elif "warping_feat_output" in output_all[0]:
        feat = output_all[0]["warping_feat_output"]
    else:
        feat = None
    return output_all, feat


def feat_key_placeholder(device):
    """np.zeros((1, DFF_FEAT_DIM, 1, 1)) of demo.py:180 / loader.py:290."""
    return torch.zeros(1, FEAT_DIM, 1, 1, device=device)
