#!/usr/bin/env python
"""Runs the REFERENCE'S OWN graph-building code on the synthetic clip and writes the committed fixtures from it (CPU, ~1 min).

`oracle/mxstub.py` lets dff_deeplab/symbols/accel_{18,34,50,101}.py + resnet_v1_101_flownet_deeplab.py build their key /
cur test symbols without MXNet and evaluates the recorded graphs with the operator functions of oracle/ops.py.  This
script drives those graphs through the chained loop of dff_deeplab/demo.py:228-250 on the same seeded weights and
3-frame clip as tests/golden/make_golden.py and

  * REQUIRES every full output of every frame (score volume, carried feature, flow, label map) to equal, bit for bit,
    what oracle/nets.py + oracle/schedule.py compute in the same process -- i.e. the oracle, written by reading the
    reference, computes exactly what the reference's own files wire up;
  * writes tests/golden/accel_<v>_128x256.npz (same arrays as make_golden.py) FROM THE REFERENCE-WIRED RUN, and
    tests/golden/reference_wired_128x256.npz: the SHA-256 of every array just written (so the tests can check, without
    /root/reference, that the committed fixtures are the files this script wrote) and the reference graphs' argument /
    auxiliary-state inventory with shapes (`sym.list_arguments()`, `sym.list_auxiliary_states()`), which the C ABI's
    `accel_param_info` must reproduce.

'dff' (Deep Feature Flow = the L branch alone, BASELINE config 2) has no symbol class in the reference: it is read off
accel_18's cur graph as the internal outputs `croped_score_output` / `warping_feat_output` (accel_18.py:172-197).

    python tests/golden/make_reference_wired.py [/root/reference]
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

sys.path.insert(0, HERE)

import make_golden  # noqa: E402
from accel_b200 import synthetic  # noqa: E402
from oracle import mxstub, ops  # noqa: E402
from oracle import schedule as oracle_schedule  # noqa: E402

H, W, FRAMES, INTERVAL, SUB = 128, 256, 3, 3, 8
DATA_INPUTS = ("data", "data_key", "feat_key")


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def build(classes, version):
    """(key symbol, cur symbol, name of the cur graph's score output) built by the reference's code."""
    cfg = mxstub.reference_config()
    inst = classes["18" if version == "dff" else version]()
    key = inst.get_key_test_symbol(cfg)
    cur = inst.get_cur_test_symbol(cfg)
    if version == "dff":
        internals = cur.get_internals()
        cur = mxstub.Group([internals["data_key"], internals["warping_feat_output"], internals["croped_score_output"]])
    score_name = "croped_score_output" if version in ("101", "dff") else "correction_output"          # demo.py:244
    return key, cur, score_name


def run_chained(key, cur, score_name, params, frames, interval):
    """dff_deeplab/demo.py:228-250 over the reference-built graphs."""
    out, feat, prev = [], None, None
    placeholder = torch.zeros(1, 2048, 1, 1)
    for idx, data in enumerate(frames):
        if prev is None:
            prev = data
        feed = dict(params)
        if idx % interval == 0:
            feed.update(data=data, data_key=prev, feat_key=placeholder)
            o = key.eval_dict(feed)
            score, feat, flow = o["croped_score_output"], o["res5c_relu_output"], None
        else:
            feed.update(data=data, data_key=prev, feat_key=feat)
            o = cur.eval_dict(feed)
            score, feat = o[score_name], o["warping_feat_output"]
            flow = cur.get_internals()["flow_grid_output"].inputs[0].eval_dict(feed)
            flow = list(flow.values())[0]
        out.append({"score": score, "feat": feat, "flow": flow, "label": ops.argmax_channel(score)[0]})
        prev = data
    return out


def _py2_prints(code):
    """`print x` -> `print(x)`, including statements that continue over several lines inside open parentheses."""
    out, lines, i = [], code.split("\n"), 0
    while i < len(lines):
        line = lines[i]
        body = line.lstrip()
        if body.startswith("print ") and not body.startswith("print ("):
            stmt = body[len("print "):]
            depth = stmt.count("(") - stmt.count(")")
            while depth > 0 and i + 1 < len(lines):
                i += 1
                stmt += "\n" + lines[i]
                depth += lines[i].count("(") - lines[i].count(")")
            out.append(line[:len(line) - len(body)] + "print(" + stmt + ")")
        else:
            out.append(line)
        i += 1
    return "\n".join(out)


def run_reference_demo_loop(ref, key, cur, version, params, frames_u8, interval, gt_labels=None):
    """Executes the frame loop of the reference's dff_deeplab/demo.py ITSELF -- from `data = []` (:165) through the
    palettised save (:256): ingest with the reference's `resize` / `transform`, `data_key` = previous frame, warm-up,
    key / cur dispatch on `idx % key_frame_interval`, `feat` carried through `im_segment`, argmax -> uint8 -- with
    stand-ins only for what is absent here: cv2 (frames come from memory), MXNet NDArray / DataBatch / Predictor
    (the predictor evaluates the reference-built graph through oracle/mxstub.py), PIL's Image (captures the arrays).
    Returns the uint8 label maps the loop handed to `Image.fromarray`.  With `gt_labels` (one (H,W) uint8 array or None
    per frame) the rest of main() runs too (:258-282: label matching by file name, `fast_hist`, `hist += curr_hist`, the
    per-frame / cumulative / final mIoU prints) and (label maps, hist, printed lines) is returned."""
    import re
    import types

    import make_reference_vectors as RV              # the ast / exec helpers and the numpy-1 shim
    RV.REF = ref

    lines = open(os.path.join(ref, "dff_deeplab/demo.py")).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.strip() == "data = []"][0]
    end = [i for i, l in enumerate(lines) if "segmentation_result.save(" in l][0]
    if gt_labels is not None:
        end = [i for i, l in enumerate(lines) if l.strip() == "print 'done'"][0] - 1
    indent = len(lines[start]) - len(lines[start].lstrip())
    code = "\n".join(l[indent:] for l in lines[start:end + 1]) + "\n"
    code = _py2_prints(code)                                                            # Python 2 print statements

    class ND:                                           # the slice of mx.nd.NDArray the loop touches
        def __init__(self, t):
            self.t = t

        @property
        def shape(self):
            return tuple(self.t.shape)

        def asnumpy(self):
            return self.t.numpy()

    class DataBatch:
        def __init__(self, data, label, pad, index, provide_data, provide_label):
            self.data, self.label, self.pad, self.index = data, label, pad, index
            self.provide_data, self.provide_label = provide_data, provide_label

    class Predictor:                                    # core/tester.py:22-35 over the reference-built graph
        def __init__(self, symbol, data_names, label_names, context=None, max_data_shapes=None, provide_data=None,
                     provide_label=None, arg_params=None, aux_params=None):
            self.symbol, self.data_names = symbol, data_names
            self.params = dict(arg_params)
            self.params.update(aux_params)

        def predict(self, batch):
            feed = dict(self.params)
            feed.update({k: v.t for k, v in zip(self.data_names, batch.data[0])})
            with torch.no_grad():
                out = self.symbol.eval_dict(feed)
            return [RV.HasKeyDict((k, ND(v)) for k, v in out.items())]

    captured = []

    class _Img:
        def __init__(self, a):
            captured.append(np.array(a, copy=True))

        def putpalette(self, p):
            assert len(p) == 768

        def save(self, path):
            pass

    frame_name = lambda i: "aachen_%06d_%06d_leftImg8bit.png" % (i // 30, i)
    label_name = lambda i: "aachen_%06d_%06d_gtFine_trainIds.png" % (i // 30, i)
    gt_by_name = {}
    if gt_labels is not None:
        gt_by_name = {label_name(i): g for i, g in enumerate(gt_labels) if g is not None}

    aux_names = set(key.list_auxiliary_states()) | set(cur.list_auxiliary_states())
    arg_params = {k: v for k, v in params.items() if k not in aux_names}
    aux_params = {k: v for k, v in params.items() if k in aux_names}
    names = [frame_name(i) for i in range(len(frames_u8))]
    by_name = dict(zip(names, [f.numpy() for f in frames_u8]))
    cv2 = types.SimpleNamespace(IMREAD_COLOR=1, IMREAD_IGNORE_ORIENTATION=128, INTER_LINEAR=1,
                                imread=lambda name, flags: by_name[name],
                                resize=lambda im, a, b, fx, fy, interpolation: (im if fx == fy == 1.0 else None))
    mx = types.SimpleNamespace(
        nd=types.SimpleNamespace(array=lambda a: ND(torch.from_numpy(np.asarray(a, dtype=np.float32)))),      # float64 -> float32
        io=types.SimpleNamespace(DataBatch=DataBatch), gpu=lambda i: ("gpu", i),
        ndarray=types.SimpleNamespace(argmax=lambda x, axis: ND(torch.from_numpy(ops.argmax_channel(x.t).astype(np.float32)))))
    h, w = frames_u8[0].shape[:2]
    ns_cfg = types.SimpleNamespace
    config = ns_cfg(SCALES=[(min(h, w), max(h, w))], network=ns_cfg(IMAGE_STRIDE=0, PIXEL_MEANS=np.array([103.06, 115.90, 123.15]),
                                                                   DFF_FEAT_DIM=2048), TEST=ns_cfg(NMS=0.3))
    img_ns = {"np": RV.np1, "cv2": cv2}
    ns = {
        "np": RV.np1, "mx": mx, "cv2": cv2, "xrange": range, "config": config, "image_names": names,
        "os": types.SimpleNamespace(path=types.SimpleNamespace(exists=lambda p: True, split=os.path.split)),
        "resize": RV._exec_function("lib/utils/image.py", "resize", img_ns),
        "transform": RV._exec_function("lib/utils/image.py", "transform", img_ns),
        "im_segment": RV._exec_function("dff_deeplab/core/tester.py", "im_segment", {}),
        "getpallete": RV._exec_function("dff_deeplab/demo.py", "getpallete", {"np": RV.np1}),
        "load_param": lambda prefix, epoch, process=False: ((arg_params, aux_params) if prefix.endswith("m1") else ({}, {})),
        "cur_path": "/nowhere/", "model1": "m1", "model2": "m2", "Predictor": Predictor, "key_sym": key, "cur_sym": cur,
        "gpu_nms_wrapper": lambda thresh, dev: None, "key_frame_interval": interval, "num_classes": 19,
        "version": "101" if version == "dff" else version,    # DFF's score output is named like Accel-101's (demo.py:244)
        "tic": lambda: None, "toc": lambda: 0.0, "output_dir": "/nowhere",
        "Image": types.SimpleNamespace(fromarray=_Img, open=lambda name: gt_by_name[name]),
        "label_files": sorted(gt_by_name),
        "fast_hist": RV._exec_function("dff_deeplab/demo.py", "fast_hist", {"np": RV.np1}),
        "per_class_iu": RV._exec_function("dff_deeplab/demo.py", "per_class_iu", {"np": RV.np1}),
    }
    import contextlib
    import io
    printed = io.StringIO()
    with contextlib.redirect_stdout(printed), np.errstate(divide="ignore", invalid="ignore"):
        exec(compile(code, "dff_deeplab/demo.py:%d-%d" % (start + 1, end + 1), "exec"), ns)
    if gt_labels is None:
        return captured
    return captured, np.asarray(ns["hist"]), printed.getvalue().splitlines()


def inventory(key, cur, params):
    args, auxs = [], []
    for s in (key, cur):
        args += [n for n in s.list_arguments() if n not in DATA_INPUTS and n not in args]
        auxs += [n for n in s.list_auxiliary_states() if n not in auxs]
    shapes = lambda names: ["x".join(str(int(x)) for x in params[n].shape) for n in names]
    return args, shapes(args), auxs, shapes(auxs)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    torch.set_num_threads(os.cpu_count() or 1)
    classes = mxstub.load_reference_symbols(ref)
    out = {"height": H, "width": W, "interval": INTERVAL, "sub": SUB}
    for version in ("dff", "18", "34", "50", "101"):
        params = synthetic.make_params(version)
        frames = synthetic.make_frames(FRAMES, H, W)
        key, cur, score_name = build(classes, version)
        with torch.no_grad():
            res = run_chained(key, cur, score_name, params, frames, INTERVAL)
            orc = oracle_schedule.run(params, version, frames, INTERVAL, "chained", keep=("label", "score", "feat", "flow"))
        for i, (a, b) in enumerate(zip(res, orc)):
            for k in ("score", "feat", "flow"):
                if a[k] is None and k not in b:
                    continue
                if not torch.equal(a[k], b[k]):
                    raise SystemExit("%s frame %d: reference-wired %s differs from the oracle (max abs %.3e)"
                                     % (version, i, k, (a[k] - b[k]).abs().max().item()))
            if not np.array_equal(np.asarray(a["label"]), np.asarray(b["label"])):
                raise SystemExit("%s frame %d: label maps differ" % (version, i))
        # the reference's own demo.py frame loop, executed over the same graphs from uint8 BGR frames, ends in the same
        # uint8 label maps
        frames_u8 = synthetic.make_frames_u8(FRAMES, H, W)
        for f8, f in zip(frames_u8, frames):
            assert torch.equal(synthetic.transform(f8), f)
        demo_labels = run_reference_demo_loop(ref, key, cur, version, params, frames_u8, INTERVAL)
        assert len(demo_labels) == FRAMES
        for i, (lab, r) in enumerate(zip(demo_labels, res)):
            if lab.dtype != np.uint8 or not np.array_equal(lab, np.asarray(r["label"])):
                raise SystemExit("%s frame %d: demo.py's loop produced a different label map" % (version, i))
        out["%s_demo_loop_frames" % version] = np.int64(len(demo_labels))
        g = make_golden.pack(res)
        np.savez_compressed(os.path.join(HERE, "accel_%s_%dx%d.npz" % (version, H, W)), **g)
        for k, a in g.items():
            out["%s_sha_%s" % (version, k)] = sha(np.asarray(a))
        args, arg_shapes, auxs, aux_shapes = inventory(key, cur, params)
        out["%s_args" % version], out["%s_arg_shapes" % version] = np.array(args), np.array(arg_shapes)
        out["%s_auxs" % version], out["%s_aux_shapes" % version] = np.array(auxs), np.array(aux_shapes)
        out["%s_key_outputs" % version] = np.array(key.list_outputs())
        out["%s_cur_outputs" % version] = np.array(cur.list_outputs())
        print("%-3s reference-wired == oracle bit for bit on %d frames; %d arguments, %d auxiliary states; outputs %s / %s"
              % (version, len(res), len(args), len(auxs), key.list_outputs(), cur.list_outputs()))
    path = os.path.join(HERE, "reference_wired_%dx%d.npz" % (H, W))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
