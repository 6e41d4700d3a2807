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
