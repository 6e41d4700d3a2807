"""Graph wiring PINNED against the reference's own code (CPU).

tests/golden/make_reference_wired.py executes dff_deeplab/symbols/accel_{18,34,50,101}.py +
resnet_v1_101_flownet_deeplab.py themselves (on oracle/mxstub.py, a stand-in for `mxnet.symbol` that evaluates with
oracle/ops.py), checks the result against oracle/nets.py bit for bit, and writes the golden fixtures from that run.  Here:

  * the committed accel_<v>_128x256.npz fixtures ARE the files that script wrote (SHA-256 of every array) -- so the GPU
    golden tests (tests/test_golden.py) compare the CUDA path with outputs of the reference's own graph code;
  * the C ABI enumerates exactly the arguments + auxiliary states of the reference's graphs, with the same shapes;
  * output names are the reference's;
  * where /root/reference is present (the build container; not the GPU box) the whole thing is re-executed live.

What this does NOT pin is the arithmetic inside each MXNet operator ([MXNet-ext] in oracle/ops.py): MXNet itself is
not available."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
import torch

from accel_b200 import _lib, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"
VERSIONS = ["dff", "18", "34", "50", "101"]
W = np.load(os.path.join(GOLDEN, "reference_wired_128x256.npz"))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("version", VERSIONS)
def test_committed_fixtures_are_the_reference_wired_run(version):
    g = np.load(os.path.join(GOLDEN, "accel_%s_128x256.npz" % version))
    keys = [k[len(version) + 5:] for k in W.files if k.startswith(version + "_sha_")]
    assert sorted(keys) == sorted(g.files) and len(keys) >= 15
    for k in keys:
        assert _sha(g[k]) == str(W["%s_sha_%s" % (version, k)]), k


@pytest.mark.parametrize("version", VERSIONS)
def test_abi_parameters_are_the_reference_graphs_arguments(lib, version):
    cfg = _lib.AccelConfig(_lib.VERSION_CODE[version], 128, 256, 19, 0, 0)
    h = C.c_void_p()
    assert lib.accel_create(C.byref(cfg), C.byref(h)) == 0, lib.accel_last_error(None)
    got = {}
    name, shape, ndim = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
    for i in range(lib.accel_param_count(h)):
        assert lib.accel_param_info(h, i, C.byref(name), shape, C.byref(ndim)) == 0
        got[name.value.decode()] = "x".join(str(shape[j]) for j in range(ndim.value))
    lib.accel_destroy(h)
    want = dict(zip([str(n) for n in W[version + "_args"]], [str(s) for s in W[version + "_arg_shapes"]]))
    want.update(zip([str(n) for n in W[version + "_auxs"]], [str(s) for s in W[version + "_aux_shapes"]]))
    # the reference's graphs also carry nodes whose output nobody reads on this path (get_flownet's
    # `Convolution5_scale`, ...flownet_deeplab.py:1805-1807): the checkpoint has them, the hot path does not need them
    dead = {k for k in want if k.startswith("Convolution5_scale")}
    missing = set(want) - dead - set(got)
    extra = set(got) - set(want)
    assert not missing and not extra, (sorted(missing)[:5], sorted(extra)[:5])
    for k in got:
        assert got[k] == want[k], (k, got[k], want[k])


def test_output_names_are_the_reference_graphs():
    for v in VERSIONS:
        assert [str(x) for x in W[v + "_key_outputs"]] == ["data_key", "feat_key", "res5c_relu_output", "croped_score_output"]
        score = "croped_score_output" if v in ("101", "dff") else "correction_output"
        assert [str(x) for x in W[v + "_cur_outputs"]] == ["data_key", "warping_feat_output", score]


def test_synthetic_parameters_cover_the_reference_inventory():
    """accel_b200/synthetic.make_params (the weights every parity test and the bench use) names exactly what the
    reference's graphs list, with the shapes the graphs imply."""
    for v in VERSIONS:
        p = synthetic.make_params(v)
        names = [str(n) for n in W[v + "_args"]] + [str(n) for n in W[v + "_auxs"]]
        shapes = [str(s) for s in W[v + "_arg_shapes"]] + [str(s) for s in W[v + "_aux_shapes"]]
        for n, s in zip(names, shapes):
            assert n in p and "x".join(str(int(x)) for x in p[n].shape) == s, n


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dff_deeplab", "symbols")), reason="reference tree not present")
@pytest.mark.parametrize("version", ["dff", "18", "101"])
def test_reference_symbol_files_executed_live_equal_oracle(version):
    import sys
    sys.path.insert(0, GOLDEN)
    import make_reference_wired as M
    from oracle import mxstub
    from oracle import schedule as oracle_schedule
    classes = mxstub.load_reference_symbols(REF)
    params = synthetic.make_params(version)
    frames = synthetic.make_frames(3, 128, 256)
    key, cur, score_name = M.build(classes, version)
    with torch.no_grad():
        res = M.run_chained(key, cur, score_name, params, frames, 3)
        orc = oracle_schedule.run(params, version, frames, 3, "chained", keep=("label", "score", "feat", "flow"))
    for a, b in zip(res, orc):
        assert torch.equal(a["score"], b["score"]) and torch.equal(a["feat"], b["feat"])
        assert np.array_equal(np.asarray(a["label"]), np.asarray(b["label"]))
        if a["flow"] is not None:
            assert torch.equal(a["flow"], b["flow"])


def test_fixture_records_the_demo_loop_check():
    """make_reference_wired.py also executed the frame loop of the reference's dff_deeplab/demo.py (:165-256) over the
    same graphs from uint8 frames and required its uint8 label maps to equal the fixtures'."""
    for v in VERSIONS:
        assert int(W[v + "_demo_loop_frames"]) == 3


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dff_deeplab", "symbols")), reason="reference tree not present")
def test_reference_demo_loop_executed_live_equals_scheduler_oracle():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_reference_wired as M
    from oracle import mxstub
    from oracle import schedule as oracle_schedule
    version, interval = "18", 2
    classes = mxstub.load_reference_symbols(REF)
    params = synthetic.make_params(version)
    frames_u8 = synthetic.make_frames_u8(5, 128, 256, stream=1)
    key, cur, _ = M.build(classes, version)
    rng = np.random.RandomState(0)
    gts = [None, rng.randint(0, 19, (128, 256)).astype(np.uint8), None, rng.randint(0, 19, (128, 256)).astype(np.uint8), None]
    gts[3][::3] = 255                                                  # Cityscapes ignore label
    labels, hist, printed = M.run_reference_demo_loop(REF, key, cur, version, params, frames_u8, interval, gt_labels=gts)
    with torch.no_grad():
        orc = oracle_schedule.run(params, version, [synthetic.transform(f) for f in frames_u8], interval, "chained")
    assert len(labels) == 5
    for lab, r in zip(labels, orc):
        assert lab.dtype == np.uint8 and np.array_equal(lab, np.asarray(r["label"]))
    # the accuracy tail of main() (:258-282): label matching, fast_hist accumulation, nanmean / round of the mIoU prints
    from oracle import io as oio
    mine = sum(oio.fast_hist(np.asarray(orc[i]["label"]).flatten(), gts[i].flatten(), 19) for i in (1, 3))
    assert np.array_equal(mine, hist)
    cum = [l for l in printed if l.startswith("(cum) mIoU")]
    final = [l for l in printed if l.startswith("===> final mIoU")]
    assert len(cum) == 2 and len(final) == 1
    assert final[0] == "===> final mIoU {mIoU:.3f}".format(mIoU=oio.mean_iou(mine))
    first = oio.fast_hist(np.asarray(orc[1]["label"]).flatten(), gts[1].flatten(), 19)
    assert cum[0] == "(cum) mIoU {mIoU:.3f}".format(mIoU=oio.mean_iou(first))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dff_deeplab", "symbols")), reason="reference tree not present")
def test_operator_defaults_the_reference_graphs_rely_on():
    """The stub fills MXNet's documented defaults where the reference omits an argument.  On the inference graphs that
    happens only for: Convolution no_bias=False / stride 1 / pad 0, Deconvolution pad 0, Concat dim=1 and
    pooling_convention='valid' on the two unnamed max pools of the R18/R34 trunks -- never for BatchNorm's eps /
    fix_gamma, Deconvolution's no_bias, LeakyReLU's slope or anything of DeformableConvolution."""
    from oracle import mxstub
    classes = mxstub.load_reference_symbols(REF)
    cfg = mxstub.reference_config()
    relied = set()
    for v in ("18", "34", "50", "101"):
        inst = classes[v]()
        for sym in (inst.get_key_test_symbol(cfg), inst.get_cur_test_symbol(cfg)):
            for n in sym._walk():
                a = n.attrs
                need = {"BatchNorm": ("eps", "fix_gamma"), "Convolution": ("no_bias", "stride", "pad", "kernel", "num_filter"),
                        "Deconvolution": ("no_bias", "pad", "stride", "kernel", "num_filter"), "Pooling": ("pooling_convention", "kernel", "stride", "pool_type"),
                        "DeformableConvolution": ("no_bias", "stride", "pad", "dilate", "kernel", "num_filter", "num_deformable_group"),
                        "LeakyReLU": ("slope", "act_type"), "Concat": ("dim",), "Crop": ("offset",),
                        "GridGenerator": ("transform_type",), "Activation": ("act_type",)}.get(n.op, ())
                relied |= {(n.op, k) for k in need if k not in a}
    assert relied == {("Convolution", "no_bias"), ("Convolution", "stride"), ("Convolution", "pad"), ("Deconvolution", "pad"),
                      ("Concat", "dim"), ("Pooling", "pooling_convention")}, relied

