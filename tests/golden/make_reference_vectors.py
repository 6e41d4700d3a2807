#!/usr/bin/env python
"""Golden vectors produced by EXECUTING the reference's own host-side Python (CPU, seconds).

MXNet, cv2 and Python 2 are absent, so the reference's modules cannot be imported and the network arithmetic (which
lives in MXNet) cannot be run -- but the pure numpy / pure Python functions around the hot path can.  This script reads
them from /root/reference AT GENERATION TIME with `ast`, executes the extracted definitions unmodified in a namespace
that only supplies numpy (plus the aliases Python 3 / numpy 2 removed: `xrange`, `np.int`, `dict.has_key`), feeds them
seeded inputs and stores inputs + outputs in tests/golden/reference_host_vectors.npz.  No reference source is copied
into the repository; the tests (tests/test_reference_vectors.py) only read the .npz.

    python tests/golden/make_reference_vectors.py [/root/reference]

Functions executed (file:line in the reference):
  transform                    lib/utils/image.py:224-235
  fast_hist, per_class_iu      dff_deeplab/demo.py:50-56
  getpallete                   dff_deeplab/demo.py:58-104
  im_segment                   dff_deeplab/core/tester.py:158-171
  TestLoader.next / get_batch  dff_deeplab/core/loader.py:259-303   (key_frame_flag stream, data_key bookkeeping)
  greedy video -> GPU split    dff_rfcn/function/test_rcnn.py:60-67
  config + update_config       dff_deeplab/config/config.py on experiments/dff_deeplab/cfgs/dff_deeplab_vid_demo.yaml
  load_param, load_param_multi lib/utils/load_model.py:4-116          (whole module, `mxnet` replaced by a stub whose
                                                                      nd.load returns the dict the test also saves)
"""
import ast
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"                    # main() takes an override from the command line


class _NP(types.ModuleType):
    """numpy with the numpy-1 alias the reference uses (np.int)."""

    def __getattr__(self, name):
        return getattr(numpy, name)


np1 = _NP("np")
np1.int = int


class HasKeyDict(dict):
    def has_key(self, k):                  # Python 2 dict API used by tester.py
        return k in self


def _tree(rel):
    path = os.path.join(REF, rel)
    src = open(path).read()
    try:
        return src, ast.parse(src)
    except SyntaxError:
        # Python 2 print statements elsewhere in the file: parse only the functions we need, by line slicing
        return src, None


def _function_source(rel, name, cls=None):
    """Source text of def `name` (optionally a method of class `cls`), dedented."""
    src, tree = _tree(rel)
    lines = src.splitlines()
    if tree is not None:
        scope = tree.body
        if cls is not None:
            scope = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
        node = [n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name][0]
        seg = lines[node.lineno - 1:node.end_lineno]
    else:
        start = [i for i, l in enumerate(lines) if l.lstrip().startswith("def %s(" % name)][0]
        indent = len(lines[start]) - len(lines[start].lstrip())
        end = start + 1
        while end < len(lines) and (not lines[end].strip() or len(lines[end]) - len(lines[end].lstrip()) > indent):
            end += 1
        seg = lines[start:end]
    indent = len(seg[0]) - len(seg[0].lstrip())
    return "\n".join(l[indent:] for l in seg) + "\n"


def _exec_function(rel, name, ns, cls=None):
    code = _function_source(rel, name, cls)
    exec(compile(code, "%s:%s" % (rel, name), "exec"), ns)
    return ns[name]


def main():
    global REF
    if len(sys.argv) > 1:
        REF = sys.argv[1]
    out = {}
    rng = numpy.random.RandomState(20260101)

    # ---- lib/utils/image.py: transform -------------------------------------------------------------------------
    ns = {"np": np1}
    transform = _exec_function("lib/utils/image.py", "transform", ns)
    means = numpy.array([103.06, 115.90, 123.15])                # experiments/dff_deeplab/cfgs/*.yaml PIXEL_MEANS (BGR)
    for i, (h, w) in enumerate([(4, 6), (17, 33), (64, 128)]):
        im = rng.randint(0, 256, size=(h, w, 3)).astype(numpy.uint8)
        out["transform_in_%d" % i] = im
        out["transform_out_%d" % i] = transform(im, means)       # float64 (1,3,h,w); mx.nd.array then casts to float32
    out["transform_means"] = means
    ramp = numpy.repeat(numpy.arange(256, dtype=numpy.uint8)[None, :, None], 3, axis=2)
    out["transform_in_ramp"] = ramp
    out["transform_out_ramp"] = transform(ramp, means)

    # ---- dff_deeplab/demo.py: fast_hist, per_class_iu, getpallete ------------------------------------------------
    ns = {"np": np1}
    fast_hist = _exec_function("dff_deeplab/demo.py", "fast_hist", ns)
    per_class_iu = _exec_function("dff_deeplab/demo.py", "per_class_iu", ns)
    getpallete = _exec_function("dff_deeplab/demo.py", "getpallete", ns)
    n = 19
    pred = rng.randint(0, n, size=(48, 80)).astype(numpy.uint8)
    label = rng.randint(0, n, size=(48, 80)).astype(numpy.uint8)
    label[rng.rand(48, 80) < 0.25] = 255                         # Cityscapes ignore label
    label[:, :3] = 7                                             # one class over-represented
    hist = fast_hist(pred.flatten(), label.flatten(), n)
    out["hist_pred"], out["hist_label"], out["hist_out"] = pred, label, hist
    with numpy.errstate(divide="ignore", invalid="ignore"):
        out["iu_out"] = per_class_iu(hist)
        sparse = numpy.zeros((n, n), dtype=numpy.int64)
        sparse[2, 2], sparse[2, 5], sparse[5, 5] = 10, 3, 4      # classes that never occur: 0/0 -> nan, as in the reference
        out["iu_sparse_in"], out["iu_sparse_out"] = sparse, per_class_iu(sparse)
    out["pallete_256"] = getpallete(256)
    out["pallete_19"] = getpallete(35)

    # ---- dff_deeplab/core/tester.py: im_segment --------------------------------------------------------------------
    ns = {}
    im_segment = _exec_function("dff_deeplab/core/tester.py", "im_segment", ns)

    class FakePredictor:
        def __init__(self, keys):
            self.keys = keys

        def predict(self, batch):
            return [HasKeyDict((k, "%s@%s" % (k, batch)) for k in self.keys)]

    picks = []
    for keys in (["data_key", "feat_key", "res5c_relu_output", "croped_score_output"],      # accel_18.py:157
                 ["data_key", "warping_feat_output", "correction_output"],                 # accel_18.py:237
                 ["data_key", "warping_feat_output", "croped_score_output"],               # accel_101.py:191
                 ["croped_score_output"]):
        output_all, feat = im_segment(FakePredictor(keys), "b")
        picks.append("" if feat is None else feat.split("@")[0])
        assert list(output_all[0].keys()) == keys
    out["im_segment_feat_key"] = numpy.array(picks)

    # ---- dff_deeplab/core/loader.py: TestLoader.next / get_batch ------------------------------------------------------
    class _MXIO:
        @staticmethod
        def DataBatch(**kw):
            return kw

    class _MXND:
        @staticmethod
        def array(a):
            return a

    mx = types.SimpleNamespace(io=_MXIO, nd=_MXND)

    def get_rpn_testbatch(roidbs, cfg):                          # stands in for image loading: the "frame" is its name
        name = roidbs[0]["image"]
        return [{"data": numpy.array([name]), "im_info": numpy.array([[1.0, 1.0, 1.0]])}], {}, [numpy.array([[1.0, 1.0, 1.0]])]

    ns = {"np": np1, "mx": mx, "get_rpn_testbatch": get_rpn_testbatch, "xrange": range}
    methods = {m: _exec_function("dff_deeplab/core/loader.py", m, ns, cls="TestLoader")
               for m in ("iter_next", "next", "get_batch", "getindex", "getpad")}
    Loader = type("Loader", (object,), methods)
    seg_lens = [7, 3, 12, 1, 5]
    rows = []
    for interval in (1, 2, 3, 5, 10):
        ld = Loader()
        ld.cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(KEY_FRAME_INTERVAL=interval),
                                       network=types.SimpleNamespace(DFF_FEAT_DIM=2048))
        ld.roidb = [{"pattern": "v%d/%%06d" % v, "frame_seg_len": L} for v, L in enumerate(seg_lens)]
        ld.size = sum(seg_lens)
        ld.batch_size = 1
        ld.cur = ld.cur_roidb_index = ld.cur_frameid = ld.key_frameid = 0          # loader.py reset()
        ld.cur_seg_len = 0
        ld.data_name = ["data", "im_info", "data_key", "feat_key"]
        ld.label = ld.provide_data = ld.provide_label = None
        while ld.iter_next():
            im_info, flag, batch = ld.next()
            data = batch["data"][0]
            rows.append((interval, int(flag), str(data[0][0]), str(data[2][0]), tuple(data[3].shape) == (1, 2048, 1, 1)))
    out["loader_seg_lens"] = numpy.array(seg_lens)
    out["loader_interval"] = numpy.array([r[0] for r in rows])
    out["loader_flag"] = numpy.array([r[1] for r in rows])
    out["loader_frame"] = numpy.array([r[2] for r in rows])
    out["loader_data_key"] = numpy.array([r[3] for r in rows])
    assert all(r[4] for r in rows)

    # ---- dff_rfcn/function/test_rcnn.py: greedy split of videos over GPUs ---------------------------------------------
    src, tree = _tree("dff_rfcn/function/test_rcnn.py")
    lines = src.splitlines()
    start = [i for i, l in enumerate(lines) if l.strip() == "gpu_num = len(ctx)"][0]
    end = [i for i, l in enumerate(lines) if "roidbs_seg_lens[gpu_id] += x['frame_seg_len']" in l][0]
    seg = lines[start:end + 1]
    indent = len(seg[0]) - len(seg[0].lstrip())
    code = "\n".join(l[indent:] for l in seg) + "\n"
    for j, (gpus, lens) in enumerate([(2, [5, 3, 9, 1, 1, 7]), (4, [10, 10, 3, 2, 8, 8, 1, 30, 4]), (8, list(range(1, 20))),
                                      (3, [4, 4, 4, 4])]):
        ns = {"np": np1, "ctx": list(range(gpus)), "roidb": [{"frame_seg_len": L, "id": i} for i, L in enumerate(lens)]}
        exec(compile(code, "test_rcnn.py:split", "exec"), ns)
        out["shard_lens_%d" % j] = numpy.array(lens)
        out["shard_gpus_%d" % j] = numpy.array(gpus)
        assign = numpy.full(len(lens), -1)
        for g, lst in enumerate(ns["roidbs"]):
            for x in lst:
                assign[x["id"]] = g
        out["shard_assign_%d" % j] = assign

    # ---- lib/utils/load_model.py: key handling of load_param / load_param_multi ---------------------------------------
    files = {
        "A-0000.params": {"arg:fc6_weight": 1.0, "arg:fc6_bias": 2.0, "aux:bn0_moving_mean": 3.0, "arg:score_weight_test": 4.0,
                          "arg:score_weight": 5.0, "aux:bn_test_moving_var": 6.0, "arg:18_conv0_weight": 7.0, "junk:x": 8.0},
        "B-0000.params": {"arg:fc6_weight": 11.0, "aux:bn0_moving_mean": 12.0, "arg:res5c_branch2c_weight": 13.0,
                          "aux:bn5c_branch2c_moving_var": 14.0, "arg:corr_bias_test": 15.0},
        "A-0003.params": {"arg:18_fc6_weight": 21.0, "arg:fc6_weight": 22.0, "aux:18_bn0_moving_var": 23.0, "aux:bn0_moving_var": 24.0},
    }
    fake_mx = types.ModuleType("mxnet")
    fake_mx.nd = types.SimpleNamespace(load=lambda fname: dict(files[os.path.basename(fname)]))
    fake_mx.cpu = lambda: "cpu"
    sys.modules["mxnet"] = fake_mx
    ns = {"__name__": "ref_load_model"}
    exec(compile(open(os.path.join(REF, "lib/utils/load_model.py")).read(), "lib/utils/load_model.py", "exec"), ns)
    del sys.modules["mxnet"]
    calls = {
        "plain": lambda: ns["load_param"]("/x/A", 0),
        "process": lambda: ns["load_param"]("/x/A", 0, process=True),
        "argprefix": lambda: ns["load_param"]("/x/A", 3, process=True, argprefix="18_"),
        "multi": lambda: ns["load_param_multi"]("/x/A", "/x/B", 0, process=True),
    }
    out["loadparam_files"] = numpy.array(sorted(files))
    for fname, d in files.items():
        out["loadparam_file_keys_" + fname] = numpy.array(list(d))
        out["loadparam_file_vals_" + fname] = numpy.array(list(d.values()))
    for tag, fn in calls.items():
        arg, aux = fn()
        out["loadparam_%s_arg_keys" % tag] = numpy.array(sorted(arg))
        out["loadparam_%s_arg_vals" % tag] = numpy.array([arg[k] for k in sorted(arg)])
        out["loadparam_%s_aux_keys" % tag] = numpy.array(sorted(aux))
        out["loadparam_%s_aux_vals" % tag] = numpy.array([aux[k] for k in sorted(aux)])

    # ---- dff_deeplab/config/config.py + experiments/dff_deeplab/cfgs/dff_deeplab_vid_demo.yaml ----------------------------
    class EasyDict(dict):                              # easydict.EasyDict: attribute access, nested dicts converted
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            super().__setitem__(k, EasyDict(v) if isinstance(v, dict) and not isinstance(v, EasyDict) else v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    import yaml as real_yaml
    fake_easydict = types.ModuleType("easydict")
    fake_easydict.EasyDict = EasyDict
    fake_yaml = types.ModuleType("yaml")
    fake_yaml.load = lambda f: real_yaml.safe_load(f)          # yaml.load(f) without a Loader is PyYAML < 6 API
    saved = {k: sys.modules.get(k) for k in ("easydict", "yaml")}
    sys.modules["easydict"], sys.modules["yaml"] = fake_easydict, fake_yaml
    try:
        ns = {"__name__": "ref_config"}
        exec(compile(open(os.path.join(REF, "dff_deeplab/config/config.py")).read(), "dff_deeplab/config/config.py", "exec"), ns)
        ns["update_config"](os.path.join(REF, "experiments/dff_deeplab/cfgs/dff_deeplab_vid_demo.yaml"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    cfg = ns["config"]
    out["config_SCALES"] = numpy.array(cfg.SCALES[0])
    out["config_PIXEL_MEANS"] = numpy.array(cfg.network.PIXEL_MEANS, dtype=numpy.float64)
    out["config_IMAGE_STRIDE"] = numpy.int64(cfg.network.IMAGE_STRIDE)
    out["config_DFF_FEAT_DIM"] = numpy.int64(cfg.network.DFF_FEAT_DIM)
    out["config_NUM_CLASSES"] = numpy.int64(cfg.dataset.NUM_CLASSES)
    out["config_KEY_FRAME_INTERVAL"] = numpy.int64(cfg.TEST.KEY_FRAME_INTERVAL)

    path = os.path.join(HERE, "reference_host_vectors.npz")
    numpy.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %d bytes)" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
