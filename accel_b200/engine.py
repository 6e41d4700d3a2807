"""Engine: one AccelHandle (weights + plans for one Accel version and frame size on one GPU).

PyTorch is only the device-memory / stream plumbing here: tensors go to the C ABI as raw pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .netspec import FEAT_DIM, NUM_CLASSES


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _check_f32_cuda(name, t, shape, device=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA float32 tensor" % name)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    if device is not None and t.device != device:
        raise ValueError("%s lives on %s, the engine on %s" % (name, t.device, device))
    if t.data_ptr() % 16:
        raise ValueError("%s must be 16-byte aligned (got a view at an odd offset)" % name)


def _check_u8_cuda(name, t, shape, device=None):
    """Label maps are written with 4-byte (uchar4) stores: uint8, contiguous, exact shape, 4-byte aligned."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA uint8 tensor" % name)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    if device is not None and t.device != device:
        raise ValueError("%s lives on %s, the engine on %s" % (name, t.device, device))
    if t.data_ptr() % 4:
        raise ValueError("%s must be 4-byte aligned (got a view at an odd offset)" % name)


class Engine:
    """Owns the device copies of the weights and the key/cur plans; the caller owns all I/O tensors
    (SURVEY.md section 8b "Ownership")."""

    def __init__(self, version, height, width, params=None, device=0, num_classes=NUM_CLASSES, flags=0, interval=None):
        """interval: also build the whole-interval plan for that many frames (accel_plan_interval; chained schedule)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("accel_b200: no CUDA device visible -- this path has no CPU fallback")
        self.version = str(version)
        if self.version not in _lib.VERSION_CODE:
            raise ValueError("unknown Accel version %r" % (version,))
        self.height, self.width, self.num_classes, self.device = int(height), int(width), int(num_classes), int(device)
        cfg = _lib.AccelConfig(_lib.VERSION_CODE[self.version], self.height, self.width, self.num_classes, self.device,
                               int(flags))
        h = C.c_void_p()
        if self.lib.accel_create(C.byref(cfg), C.byref(h)) != 0:
            raise RuntimeError("accel_create failed: %s" % self.lib.accel_last_error(None).decode())
        self._h = h
        self.flags = int(flags)
        self.torch_device = torch.device("cuda", self.device)
        self.interval = 0
        if interval is not None and int(interval) > 1:
            self.plan_interval(int(interval))
        if params is not None:
            self.set_params(params)

    def plan_interval(self, interval):
        """accel_plan_interval: must precede finalize (set_params(..., finalize=True))."""
        if self.lib.accel_plan_interval(self._h, int(interval)) != 0:
            raise RuntimeError("accel_plan_interval failed: %s" % self._err())
        self.interval = int(interval)

    @property
    def supports_interval(self):
        return self.interval > 1

    # ---- parameters -----------------------------------------------------------------------------
    def param_spec(self):
        out = {}
        name, shape, ndim = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
        for i in range(self.lib.accel_param_count(self._h)):
            self.lib.accel_param_info(self._h, i, C.byref(name), shape, C.byref(ndim))
            out[name.value.decode()] = tuple(shape[j] for j in range(ndim.value))
        return out

    def set_params(self, params, finalize=True):
        """params: {reference parameter name: array-like fp32}; arg_params and aux_params merged
        (lib/utils/load_model.py:73-93).  Unknown names are ignored the way MXNet's
        init_params(allow_extra) would; missing ones fail at finalize."""
        spec = self.param_spec()
        for name, value in params.items():
            if name not in spec:
                continue
            a = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
            a = np.ascontiguousarray(a, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            if self.lib.accel_set_param(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim) != 0:
                raise RuntimeError(self._err())
        if finalize:
            self.finalize()

    def finalize(self):
        with torch.cuda.device(self.torch_device):
            if self.lib.accel_finalize(self._h) != 0:
                raise RuntimeError("accel_finalize failed: %s" % self._err())

    # ---- forward ------------------------------------------------------------------------------------
    @property
    def feat_shape(self):
        return (1, FEAT_DIM, self.height // 16, self.width // 16)

    @property
    def g_shape(self):
        """fc6's linear part W*F that the commuted L head warps: (1, 1024, H/16, W/16)."""
        return (1, 1024, self.height // 16, self.width // 16)

    @property
    def supports_linear_head(self):
        return self.version != "101"

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def key_forward(self, data, feat_out=None, score_out=None, label_out=None, g_out=None):
        """g_out: optional (1,1024,H/16,W/16) output of fc6's linear part (accel_key_forward_lin)."""
        dv = self.torch_device
        _check_f32_cuda("data", data, (1, 3, self.height, self.width), dv)
        if feat_out is not None:
            _check_f32_cuda("feat_out", feat_out, self.feat_shape, dv)
        if score_out is not None:
            _check_f32_cuda("score_out", score_out, (1, self.num_classes, self.height, self.width), dv)
        if g_out is not None:
            _check_f32_cuda("g_out", g_out, self.g_shape, dv)
        if label_out is not None:
            _check_u8_cuda("label_out", label_out, (self.height, self.width), dv)
        with torch.cuda.device(self.torch_device):
            if g_out is not None:
                rc = self.lib.accel_key_forward_lin(self._h, _ptr(data), _ptr(feat_out), _ptr(g_out), _ptr(score_out),
                                                    _ptr(label_out), self._stream())
            else:
                rc = self.lib.accel_key_forward(self._h, _ptr(data), _ptr(feat_out), _ptr(score_out), _ptr(label_out),
                                                self._stream())
        if rc != 0:
            raise RuntimeError("accel_key_forward failed: %s" % self._err())

    def cur_forward(self, data, data_key, feat_key, feat_out=None, score_out=None, label_out=None):
        dv = self.torch_device
        _check_f32_cuda("data", data, (1, 3, self.height, self.width), dv)
        _check_f32_cuda("data_key", data_key, (1, 3, self.height, self.width), dv)
        _check_f32_cuda("feat_key", feat_key, self.feat_shape, dv)
        if feat_out is not None:
            _check_f32_cuda("feat_out", feat_out, self.feat_shape, dv)
        if score_out is not None:
            _check_f32_cuda("score_out", score_out, (1, self.num_classes, self.height, self.width), dv)
        if label_out is not None:
            _check_u8_cuda("label_out", label_out, (self.height, self.width), dv)
        with torch.cuda.device(self.torch_device):
            rc = self.lib.accel_cur_forward(self._h, _ptr(data), _ptr(data_key), _ptr(feat_key), _ptr(feat_out),
                                            _ptr(score_out), _ptr(label_out), self._stream())
        if rc != 0:
            raise RuntimeError("accel_cur_forward failed: %s" % self._err())

    def cur_forward_lin(self, data, data_key, g_key, g_out=None, score_out=None, label_out=None):
        """Cur-frame graph with the L head commuted through the warp (accel_cur_forward_lin): warps g_key = W_fc6 * F
        instead of the 2048-channel feature; DFF / Accel-18/34/50 only."""
        dv = self.torch_device
        _check_f32_cuda("data", data, (1, 3, self.height, self.width), dv)
        _check_f32_cuda("data_key", data_key, (1, 3, self.height, self.width), dv)
        _check_f32_cuda("g_key", g_key, self.g_shape, dv)
        if g_out is not None:
            _check_f32_cuda("g_out", g_out, self.g_shape, dv)
        if score_out is not None:
            _check_f32_cuda("score_out", score_out, (1, self.num_classes, self.height, self.width), dv)
        if label_out is not None:
            _check_u8_cuda("label_out", label_out, (self.height, self.width), dv)
        with torch.cuda.device(self.torch_device):
            rc = self.lib.accel_cur_forward_lin(self._h, _ptr(data), _ptr(data_key), _ptr(g_key), _ptr(g_out),
                                                _ptr(score_out), _ptr(label_out), self._stream())
        if rc != 0:
            raise RuntimeError("accel_cur_forward_lin failed: %s" % self._err())

    def rbranch_forward(self, data, score_out=None, label_out=None):
        """accel_rbranch_forward: the correction network + its head alone = plain DeepLab-<v> on one frame."""
        dv = self.torch_device
        _check_f32_cuda("data", data, (1, 3, self.height, self.width), dv)
        if score_out is not None:
            _check_f32_cuda("score_out", score_out, (1, self.num_classes, self.height, self.width), dv)
        if label_out is not None:
            _check_u8_cuda("label_out", label_out, (self.height, self.width), dv)
        with torch.cuda.device(dv):
            rc = self.lib.accel_rbranch_forward(self._h, _ptr(data), _ptr(score_out), _ptr(label_out), self._stream())
        if rc != 0:
            raise RuntimeError("accel_rbranch_forward failed: %s" % self._err())

    def interval_forward(self, frames, labels, scores=None):
        """accel_interval_forward: one whole key interval (chained schedule).  frames: `interval` fp32 (1,3,H,W) CUDA
        tensors; labels: `interval` uint8 (H,W) CUDA tensors (or one (interval,H,W) tensor); scores: None or a list of
        (1,19,H,W) tensors / None."""
        I = self.interval
        if I < 2:
            raise RuntimeError("Engine was built without interval=...: no whole-interval plan")
        dv = self.torch_device
        if len(frames) != I or len(labels) != I or (scores is not None and len(scores) != I):
            raise ValueError("interval_forward needs exactly %d frames / label maps" % I)
        vp = C.c_void_p
        fr, lb, sc = (vp * I)(), (vp * I)(), (vp * I)()
        for t in range(I):
            _check_f32_cuda("frames[%d]" % t, frames[t], (1, 3, self.height, self.width), dv)
            _check_u8_cuda("labels[%d]" % t, labels[t], (self.height, self.width), dv)
            fr[t], lb[t] = frames[t].data_ptr(), labels[t].data_ptr()
            if scores is not None and scores[t] is not None:
                _check_f32_cuda("scores[%d]" % t, scores[t], (1, self.num_classes, self.height, self.width), dv)
                sc[t] = scores[t].data_ptr()
        with torch.cuda.device(dv):
            rc = self.lib.accel_interval_forward(self._h, fr, sc if scores is not None else None, lb, self._stream())
        if rc != 0:
            raise RuntimeError("accel_interval_forward failed: %s" % self._err())

    def fetch_layer(self, plan, op_name):
        """accel_debug_fetch: the internal output of layer `op_name` of `plan` after the last forward, as an fp32
        (frames, C, H, W) CUDA tensor (per-layer parity tests, debugging)."""
        shape = (C.c_int64 * 4)()
        if self.lib.accel_debug_fetch(self._h, plan.encode(), op_name.encode(), None, shape, None) != 0:
            raise RuntimeError(self._err())
        out = torch.empty(tuple(int(x) for x in shape), device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            rc = self.lib.accel_debug_fetch(self._h, plan.encode(), op_name.encode(), _ptr(out), shape, self._stream())
        if rc != 0:
            raise RuntimeError(self._err())
        return out

    def graph_cache_stats(self):
        """(hits, misses) of the handle's CUDA-graph cache: a miss = one stream capture + instantiate."""
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.accel_graph_cache_stats(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def flownet(self, data, data_key, flow_out=None):
        dv = self.torch_device
        _check_f32_cuda("data", data, (1, 3, self.height, self.width), dv)
        _check_f32_cuda("data_key", data_key, (1, 3, self.height, self.width), dv)
        if flow_out is None:
            flow_out = torch.empty(1, 2, self.height // 16, self.width // 16, device=self.torch_device)
        else:
            _check_f32_cuda("flow_out", flow_out, (1, 2, self.height // 16, self.width // 16), dv)
        with torch.cuda.device(self.torch_device):
            rc = self.lib.accel_flownet(self._h, _ptr(data), _ptr(data_key), _ptr(flow_out), self._stream())
        if rc != 0:
            raise RuntimeError("accel_flownet failed: %s" % self._err())
        return flow_out

    # ---- introspection ------------------------------------------------------------------------------------
    def last_launch_count(self):
        return self.lib.accel_last_launch_count(self._h)

    def set_profiling(self, on):
        self.lib.accel_set_profiling(self._h, 1 if on else 0)

    def stage_times(self):
        names, ms = (C.c_char_p * 32)(), (C.c_float * 32)()
        n = self.lib.accel_stage_times(self._h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(max(n, 0))}

    def op_times(self):
        """[(layer name, ms, reference flops)] of the last profiled forward, in launch order."""
        cap = 1024
        names, ms, fl = (C.c_char_p * cap)(), (C.c_float * cap)(), (C.c_double * cap)()
        n = self.lib.accel_op_times(self._h, names, ms, fl, cap)
        return [(names[i].decode(), float(ms[i]), float(fl[i])) for i in range(max(n, 0))]

    def _err(self):
        return self.lib.accel_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self.lib.accel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- operator-level wrappers (mirror the MXNet built-ins the symbol files call) ------------------------
def warp(feat, flow, out=None):
    """GridGenerator(transform_type='warp') + BilinearSampler, accel_18.py:174-175."""
    lib = _lib.load()
    n, c, h, w = feat.shape
    assert n == 1 and tuple(flow.shape) == (1, 2, h, w)
    if out is None:
        out = torch.empty_like(feat)
    st = C.c_void_p(torch.cuda.current_stream(feat.device).cuda_stream)
    with torch.cuda.device(feat.device):
        rc = lib.accel_warp(_ptr(feat.contiguous()), _ptr(flow.contiguous()), _ptr(out), c, h, w, st)
    if rc != 0:
        raise RuntimeError("accel_warp failed (%d)" % rc)
    return out


def warp_split(feat, flow, out=None, out_hi=None, out_lo=None):
    """The warp plus the split-fp16 NHWC operand of the head's first conv (accel_18.py:174-183), one pass over `feat`:
    returns (out fp32 (1,C,H,W), hi fp16 (H,W,C), lo fp16 (H,W,C)) with hi + lo == out to fp32 precision."""
    lib = _lib.load()
    n, c, h, w = feat.shape
    assert n == 1 and tuple(flow.shape) == (1, 2, h, w) and c % 32 == 0
    if out is None:
        out = torch.empty_like(feat)
    if out_hi is None:
        out_hi = torch.empty(h, w, c, dtype=torch.float16, device=feat.device)
    if out_lo is None:
        out_lo = torch.empty(h, w, c, dtype=torch.float16, device=feat.device)
    st = C.c_void_p(torch.cuda.current_stream(feat.device).cuda_stream)
    with torch.cuda.device(feat.device):
        rc = lib.accel_warp_split(_ptr(feat.contiguous()), _ptr(flow.contiguous()), _ptr(out), _ptr(out_hi), _ptr(out_lo),
                                  c, h, w, st)
    if rc != 0:
        raise RuntimeError("accel_warp_split failed (%d)" % rc)
    return out, out_hi, out_lo


def fuse_argmax(score_a, score_b=None, corr_weight=None, corr_bias=None, want_scores=False):
    """x16 upsampling + Crop(8,8) of the low-res score map(s), 1x1 `correction` fusion, argmax."""
    lib = _lib.load()
    n, k, h, w = score_a.shape
    dev = score_a.device
    label = torch.empty(16 * h, 16 * w, dtype=torch.uint8, device=dev)
    full = torch.empty(1, k, 16 * h, 16 * w, device=dev) if want_scores else None
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cw = corr_weight.reshape(k, 2 * k).contiguous() if corr_weight is not None else None
    with torch.cuda.device(dev):
        rc = lib.accel_fuse_argmax(_ptr(score_a.contiguous()), _ptr(score_b.contiguous() if score_b is not None else None),
                                   _ptr(cw), _ptr(corr_bias), k, h, w, _ptr(label), _ptr(full), st)
    if rc != 0:
        raise RuntimeError("accel_fuse_argmax failed (%d)" % rc)
    return (label, full) if want_scores else label


def preprocess(frame_bgr_u8, out=None, pixel_means_bgr=None):
    """lib/utils/image.py:224-235 transform() on the device: (H,W,3) uint8 BGR CUDA tensor -> (1,3,H,W) fp32
    RGB minus PIXEL_MEANS (float64 subtraction rounded once to fp32, as numpy + mx.nd.array)."""
    lib = _lib.load()
    if not (frame_bgr_u8.is_cuda and frame_bgr_u8.dtype == torch.uint8 and frame_bgr_u8.is_contiguous()
            and frame_bgr_u8.dim() == 3 and frame_bgr_u8.shape[2] == 3):
        raise TypeError("frame must be a contiguous CUDA uint8 tensor of shape (H, W, 3)")
    h, w = int(frame_bgr_u8.shape[0]), int(frame_bgr_u8.shape[1])
    if out is None:
        out = torch.empty(1, 3, h, w, device=frame_bgr_u8.device)
    else:
        _check_f32_cuda("out", out, (1, 3, h, w))
    if pixel_means_bgr is None:
        from .synthetic import PIXEL_MEANS_BGR as pixel_means_bgr
    means = (C.c_double * 3)(*[float(m) for m in pixel_means_bgr])
    st = C.c_void_p(torch.cuda.current_stream(frame_bgr_u8.device).cuda_stream)
    with torch.cuda.device(frame_bgr_u8.device):
        rc = lib.accel_preprocess(_ptr(frame_bgr_u8), h, w, means, _ptr(out), st)
    if rc != 0:
        raise RuntimeError("accel_preprocess failed (%d)" % rc)
    return out


def resize_scale(height, width, target_size, max_size):
    """im_scale of lib/utils/image.py:204-210: the short side to target_size unless that pushes the long side past max_size."""
    im_size_min, im_size_max = min(height, width), max(height, width)
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


def resize(frame_bgr_u8, target_size, max_size, stride=0):
    """lib/utils/image.py:194-222 `resize` on the device: (H,W,3) uint8 BGR CUDA tensor -> (resized tensor, im_scale);
    cv2.INTER_LINEAR bit for bit (accel_resize_bgr); stride > 0 pads bottom / right with zeros to a multiple of it."""
    lib = _lib.load()
    if not (frame_bgr_u8.is_cuda and frame_bgr_u8.dtype == torch.uint8 and frame_bgr_u8.is_contiguous()
            and frame_bgr_u8.dim() == 3 and frame_bgr_u8.shape[2] == 3):
        raise TypeError("frame must be a contiguous CUDA uint8 tensor of shape (H, W, 3)")
    h, w = int(frame_bgr_u8.shape[0]), int(frame_bgr_u8.shape[1])
    im_scale = resize_scale(h, w, target_size, max_size)
    dh, dw = C.c_int(), C.c_int()
    if lib.accel_resize_size(h, w, im_scale, im_scale, C.byref(dh), C.byref(dw)) != 0:
        raise RuntimeError("accel_resize_size failed")
    out = torch.empty(dh.value, dw.value, 3, dtype=torch.uint8, device=frame_bgr_u8.device)
    st = C.c_void_p(torch.cuda.current_stream(frame_bgr_u8.device).cuda_stream)
    with torch.cuda.device(frame_bgr_u8.device):
        rc = lib.accel_resize_bgr(_ptr(frame_bgr_u8), h, w, im_scale, im_scale, _ptr(out), st)
    if rc != 0:
        raise RuntimeError("accel_resize_bgr failed (%d)" % rc)
    if stride:
        ph = int(np.ceil(out.shape[0] / float(stride)) * stride)
        pw = int(np.ceil(out.shape[1] / float(stride)) * stride)
        padded = torch.zeros(ph, pw, 3, dtype=torch.uint8, device=out.device)
        padded[:out.shape[0], :out.shape[1]] = out
        out = padded
    return out, im_scale


def confusion(pred, label, hist=None, num_classes=NUM_CLASSES):
    """fast_hist(pred, label, n) of dff_deeplab/demo.py:50-53, accumulated into `hist` (n x n int64, CUDA)."""
    lib = _lib.load()
    for name, t in (("pred", pred), ("label", label)):
        if not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
            raise TypeError("%s must be a contiguous CUDA uint8 tensor" % name)
    if pred.numel() != label.numel():
        raise ValueError("pred and label differ in size")
    if hist is None:
        hist = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=pred.device)
    elif not (hist.is_cuda and hist.dtype == torch.int64 and hist.is_contiguous()
              and tuple(hist.shape) == (num_classes, num_classes)):
        raise TypeError("hist must be a contiguous CUDA int64 tensor of shape (n, n)")
    st = C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
    with torch.cuda.device(pred.device):
        rc = lib.accel_confusion(_ptr(pred), _ptr(label), pred.numel(), num_classes, _ptr(hist), st)
    if rc != 0:
        raise RuntimeError("accel_confusion failed (%d)" % rc)
    return hist


def head(feat, fc6_weight, fc6_bias, score_weight, score_bias):
    """accel_head: ReLU(fc6(feat)) -> score at feature resolution (accel_18.py:177-191); feat (1,C,h,w) CUDA fp32, the four
    parameters host or device tensors in MXNet layout.  Returns the (1,K,h,w) low-resolution score map."""
    lib = _lib.load()
    _, cin, h, w = feat.shape
    host = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    w1, b1, w2, b2 = host(fc6_weight), host(fc6_bias), host(score_weight), host(score_bias)
    mid, k = w1.shape[0], w2.shape[0]
    if w1.shape[1] != cin or w2.shape[1] != mid or b1.shape != (mid,) or b2.shape != (k,):
        raise ValueError("accel_head: inconsistent parameter shapes")
    out = torch.empty(1, k, h, w, device=feat.device)
    err = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    with torch.cuda.device(feat.device):
        torch.cuda.synchronize()
        rc = lib.accel_head(_ptr(feat.contiguous()), cin, h, w, p(w1), p(b1), mid, p(w2), p(b2), k, _ptr(out),
                            feat.device.index or 0, err, 512)
    if rc != 0:
        raise RuntimeError("accel_head failed: %s" % err.value.decode())
    return out


def conv_layer(x, weight, kind="conv", stride=1, pad=0, dilate=1, scale=None, shift=None, act=0, residual=None,
               offset=None, deform_groups=1, engine=0):
    """One layer through the library's kernels (parity-test hook, accel_conv_layer)."""
    lib = _lib.load()
    kinds = {"conv": 0, "deconv": 1, "deform": 2, "stem": 3, "flowstem": 4}
    n, cin, hin, win = x.shape
    w = np.ascontiguousarray(weight.detach().cpu().numpy(), dtype=np.float32)
    if kind == "deconv":
        cout, k = w.shape[1], 4
        ho, wo = 2 * hin, 2 * win
    elif kind == "flowstem":
        cin, cout, k = 6, w.shape[0], 7
        ho, wo = hin // 4, win // 4
    else:
        cout, k = w.shape[0], w.shape[2]
        ho = (hin + 2 * pad - (dilate * (k - 1) + 1)) // stride + 1
        wo = (win + 2 * pad - (dilate * (k - 1) + 1)) // stride + 1
    if kind == "flowstem":
        ho, wo = hin // 4, win // 4
    out = torch.empty(1, cout, ho, wo, device=x.device)
    sc = np.ascontiguousarray(scale.detach().cpu().numpy(), dtype=np.float32) if scale is not None else None
    sh = np.ascontiguousarray(shift.detach().cpu().numpy(), dtype=np.float32) if shift is not None else None
    err = C.create_string_buffer(512)
    with torch.cuda.device(x.device):
        torch.cuda.synchronize()
        rc = lib.accel_conv_layer(kinds[kind], _ptr(x.contiguous()), cin, hin, win, w.ctypes.data_as(C.c_void_p), cout, k,
                                  stride, pad, dilate, deform_groups, _ptr(offset.contiguous() if offset is not None else None),
                                  sc.ctypes.data_as(C.c_void_p) if sc is not None else None,
                                  sh.ctypes.data_as(C.c_void_p) if sh is not None else None, act,
                                  _ptr(residual.contiguous() if residual is not None else None), engine, _ptr(out),
                                  x.device.index or 0, err, 512)
    if rc != 0:
        raise RuntimeError("accel_conv_layer failed: %s" % err.value.decode())
    return out
