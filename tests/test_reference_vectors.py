"""Host-side rows of the path PINNED against the reference's own code: tests/golden/reference_host_vectors.npz holds
inputs and outputs of the reference's pure-Python functions EXECUTED in the build container by
tests/golden/make_reference_vectors.py (transform, fast_hist, per_class_iu, getpallete, im_segment, TestLoader's
next/get_batch bookkeeping, the greedy video -> GPU split).  The oracle restatements (oracle/io.py) and the product's
host mirror (accel_b200/{scheduler,loader,predictor,visualize,synthetic}.py) must reproduce them exactly.
CPU only; the GPU kernels behind accel_preprocess / accel_confusion are checked against the same oracle functions in
tests/test_gpu_io.py."""
import os

import numpy as np
import pytest
import torch

from accel_b200 import loader, predictor, scheduler, synthetic, visualize
from oracle import io as oio

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host_vectors.npz"))


# ---------------------------------------------------------------- lib/utils/image.py:224-235 transform
@pytest.mark.parametrize("tag", ["0", "1", "2", "ramp"])
def test_transform_equals_reference(tag):
    im, want64 = G["transform_in_" + tag], G["transform_out_" + tag]
    assert want64.dtype == np.float64                                # the reference builds float64 ...
    want = want64.astype(np.float32)                                 # ... and mx.nd.array() rounds once to float32
    got = oio.transform(im, G["transform_means"])
    assert got.dtype == np.float32 and np.array_equal(got, want)
    got_t = synthetic.transform(torch.from_numpy(im)).numpy()        # what the fixtures and bench frames are built with
    assert np.array_equal(got_t, want)


# ---------------------------------------------------------------- dff_deeplab/demo.py:50-56 fast_hist / per_class_iu
def test_fast_hist_equals_reference():
    h = oio.fast_hist(G["hist_pred"].flatten(), G["hist_label"].flatten(), 19)
    assert np.array_equal(h, G["hist_out"])
    t = scheduler.confusion_matrix(torch.from_numpy(G["hist_pred"]), torch.from_numpy(G["hist_label"]), 19).numpy()
    assert np.array_equal(t, G["hist_out"])


def test_per_class_iu_equals_reference():
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(oio.per_class_iu(G["hist_out"]), G["iu_out"], equal_nan=True)
        got = oio.per_class_iu(G["iu_sparse_in"])
    assert np.array_equal(got, G["iu_sparse_out"], equal_nan=True)   # absent classes are NaN, as in the reference
    assert np.isnan(G["iu_sparse_out"]).sum() == 17


# ---------------------------------------------------------------- dff_deeplab/demo.py:58-104 getpallete
def test_palette_equals_reference():
    assert np.array_equal(visualize.getpallete(256), G["pallete_256"])
    assert np.array_equal(visualize.getpallete(35), G["pallete_19"])


# ---------------------------------------------------------------- dff_deeplab/core/tester.py:158-171 im_segment
class _FakePredictor:
    def __init__(self, keys):
        self.keys = keys

    def predict(self, batch):
        return [{k: k for k in self.keys}]


def test_im_segment_feature_choice_equals_reference():
    cases = (["data_key", "feat_key", "res5c_relu_output", "croped_score_output"],
             ["data_key", "warping_feat_output", "correction_output"],
             ["data_key", "warping_feat_output", "croped_score_output"],
             ["croped_score_output"])
    for keys, want in zip(cases, G["im_segment_feat_key"]):
        out, feat = predictor.im_segment(_FakePredictor(keys), None)
        assert list(out[0].keys()) == keys
        assert (feat or "") == str(want)


# ---------------------------------------------------------------- dff_deeplab/core/loader.py:259-303 TestLoader
class _CpuLoader(loader.TestLoader):
    """TestLoader with the GPU ingest replaced by a tag tensor (video, frame): only the bookkeeping is under test."""

    def _ingest(self, rec, frameid):
        return torch.tensor([[[[float(rec["video"]), float(frameid)]]]])


@pytest.mark.parametrize("interval", [1, 2, 3, 5, 10])
def test_loader_flags_and_data_key_equal_reference(interval):
    seg_lens = [int(x) for x in G["loader_seg_lens"]]
    m = G["loader_interval"] == interval
    want_flag, want_frame, want_key = G["loader_flag"][m], G["loader_frame"][m], G["loader_data_key"][m]
    cfg = loader.default_config(key_frame_interval=interval)
    roidb = [{"video": v, "frame_seg_len": n, "frames": None} for v, n in enumerate(seg_lens)]
    ld = _CpuLoader(roidb, cfg, device="cpu")
    name = lambda t: "v%d/%06d" % (int(t[0, 0, 0, 0]), int(t[0, 0, 0, 1]))
    got = []
    for im_info, flag, batch in ld:
        data, _, data_key, feat_key = batch.data[0]
        got.append((int(flag), name(data), name(data_key)))
        assert tuple(feat_key.shape) == (1, 2048, 1, 1)
    assert len(got) == sum(seg_lens) == len(want_flag)
    assert [g[0] for g in got] == [int(f) for f in want_flag]
    assert [g[1] for g in got] == [str(f) for f in want_frame]
    assert [g[2] for g in got] == [str(f) for f in want_key]
    # the flag stream of scheduler.key_frame_flags is the same per video
    flat = [f for n in seg_lens for f in scheduler.key_frame_flags(n, interval)]
    assert flat == [int(f) for f in want_flag]


# ---------------------------------------------------------------- dff_rfcn/function/test_rcnn.py:60-67 greedy split
@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_shard_streams_equals_reference(case):
    lens = [int(x) for x in G["shard_lens_%d" % case]]
    gpus = int(G["shard_gpus_%d" % case])
    shards = scheduler.shard_streams(lens, gpus)
    assign = np.full(len(lens), -1)
    for g, lst in enumerate(shards):
        for i in lst:
            assign[i] = g
    assert np.array_equal(assign, G["shard_assign_%d" % case])
    for g, lst in enumerate(shards):                                  # order inside a GPU's list is arrival order
        assert lst == sorted(lst)


# ---------------------------------------------------------------- lib/utils/load_model.py:4-116 load_param family
def _write_ref_files(tmp_path):
    from accel_b200 import params_io
    for fname in G["loadparam_files"]:
        keys, vals = G["loadparam_file_keys_" + str(fname)], G["loadparam_file_vals_" + str(fname)]
        params_io.nd_save(str(tmp_path / str(fname)), {str(k): np.full((2, 3), v, dtype=np.float32) for k, v in zip(keys, vals)})


@pytest.mark.parametrize("tag", ["plain", "process", "argprefix", "multi"])
def test_load_param_key_handling_equals_reference(tag, tmp_path):
    """arg:/aux: split, argprefix, `_test` renaming (arg_params only) and the override order of the two-file merge: the
    reference's load_model.py was executed on the same dictionaries (its mx.nd.load stubbed); here they go through
    the .params writer/reader first.  `junk:` keys vanish in both."""
    from accel_b200 import params_io
    _write_ref_files(tmp_path)
    A, B = str(tmp_path / "A"), str(tmp_path / "B")
    arg, aux = {"plain": lambda: params_io.load_param(A, 0),
                "process": lambda: params_io.load_param(A, 0, process=True),
                "argprefix": lambda: params_io.load_param(A, 3, process=True, argprefix="18_"),
                "multi": lambda: params_io.load_param_multi(A, B, 0, process=True)}[tag]()
    for got, kind in ((arg, "arg"), (aux, "aux")):
        want_keys = [str(k) for k in G["loadparam_%s_%s_keys" % (tag, kind)]]
        want_vals = G["loadparam_%s_%s_vals" % (tag, kind)]
        assert sorted(got) == want_keys
        for k, v in zip(want_keys, want_vals):
            a = np.asarray(got[k])
            assert a.shape == (2, 3) and np.all(a == np.float32(v))


# ---------------------------------------------------------------- dff_deeplab/config/config.py + dff_deeplab_vid_demo.yaml
def test_default_config_equals_reference_config():
    """The reference's config.py was executed on its own demo yaml (update_config); loader.default_config, the synthetic
    generator, the engine constants and bench.py's defaults carry the same values."""
    from accel_b200 import netspec
    cfg = loader.default_config()
    assert list(cfg.SCALES[0]) == [int(x) for x in G["config_SCALES"]]
    assert np.array_equal(np.asarray(cfg.network.PIXEL_MEANS, dtype=np.float64), G["config_PIXEL_MEANS"])
    assert tuple(synthetic.PIXEL_MEANS_BGR) == tuple(float(x) for x in G["config_PIXEL_MEANS"])
    assert cfg.network.IMAGE_STRIDE == int(G["config_IMAGE_STRIDE"])
    assert cfg.network.DFF_FEAT_DIM == netspec.FEAT_DIM == int(G["config_DFF_FEAT_DIM"])
    assert cfg.dataset.NUM_CLASSES == netspec.NUM_CLASSES == int(G["config_NUM_CLASSES"])
    assert cfg.TEST.KEY_FRAME_INTERVAL == int(G["config_KEY_FRAME_INTERVAL"])
