"""TestLoader-shaped iterator + segmentation pred_eval (accel_b200/loader.py) on the GPU: the flag stream is
the reference's (core/loader.py:259-295), every frame's label map equals the plain un-chained loop's, the
confusion matrix equals fast_hist over the labelled frames, and checkpoints written/read through params_io
drive the same predictors."""
import numpy as np
import pytest
import torch

from accel_b200 import loader, params_io, predictor, scheduler, synthetic
from oracle import io as oio
from oracle import schedule as oracle_schedule

pytestmark = pytest.mark.gpu
H, W, INTERVAL = 128, 256, 3


def _roidb():
    vids = [synthetic.make_frames_u8(7, H, W, stream=4), synthetic.make_frames_u8(4, H, W, stream=5)]
    g = torch.Generator().manual_seed(3)
    roidb, start = [], 0
    for v in vids:
        labels = {t: torch.randint(0, 19, (H, W), generator=g, dtype=torch.uint8).numpy() for t in (0, 2, len(v) - 1)}
        for lab in labels.values():
            lab[::4, ::4] = 255
        roidb.append({"pattern": None, "frames": v, "frame_seg_len": len(v), "frame_id": start, "labels": labels})
        start += len(v)
    return roidb


def test_loader_flags_and_pred_eval(tmp_path):
    cfg = loader.default_config(key_frame_interval=INTERVAL, scales=(H, W))
    roidb = _roidb()
    # weights travel through a .params checkpoint pair, as demo.py:192-195 loads them
    params = {k: v.numpy() for k, v in synthetic.make_params("18").items()}
    arg, aux = params_io.split_arg_aux(params)
    params_io.save_checkpoint(str(tmp_path / "accel"), 0, arg, aux)
    arg_params, aux_params = params_io.load_param(str(tmp_path / "accel"), 0, process=True)

    test_data = loader.TestLoader(roidb, cfg, device="cuda:0")
    assert test_data.size == 11 and test_data.data_name == ["data", "im_info", "data_key", "feat_key"]
    sym = predictor.accel_18()
    shapes = [[("data", (1, 3, H, W)), ("data_key", (1, 3, H, W))]]
    key_p = predictor.Predictor(sym.get_key_test_symbol(cfg), test_data.data_name, [], context=[predictor.gpu(0)],
                                max_data_shapes=shapes, arg_params=arg_params, aux_params=aux_params)
    cur_p = predictor.Predictor(sym.get_cur_test_symbol(cfg), test_data.data_name, [], context=[predictor.gpu(0)],
                                max_data_shapes=shapes, arg_params=arg_params, aux_params=aux_params)

    # flag stream == the reference's rule, video by video
    flags = [f for _, f, _ in loader.TestLoader(roidb, cfg, device="cuda:0")]
    want = oracle_schedule.key_frame_flags(7, INTERVAL) + oracle_schedule.key_frame_flags(4, INTERVAL)
    assert flags == want

    res = loader.pred_eval(0, key_p, cur_p, test_data, None, cfg, keep_labels=True)
    assert list(res["frame_ids"]) == list(range(11))

    # plain un-chained loop over each video through the same engine
    eng = key_p.engine
    dev = eng.torch_device
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    want_hist = np.zeros((19, 19), dtype=np.int64)
    k = 0
    for rec in roidb:
        state = scheduler.StreamState(eng)
        for t in range(rec["frame_seg_len"]):
            data = synthetic.transform(rec["frames"][t]).to(dev)
            scheduler.segment_frame(eng, state, data, INTERVAL, "unchained", lab)
            ref = lab.cpu().numpy()
            assert np.array_equal(res["labels"][k], ref), "frame %d" % k
            if t in rec["labels"]:
                want_hist += oio.fast_hist(ref.flatten(), rec["labels"][t].flatten(), 19)
            k += 1
    assert np.array_equal(res["hist"], want_hist)
    merged = loader.pred_eval_multiprocess(1, [key_p], [cur_p], [loader.TestLoader(roidb, cfg, device="cuda:0")], None, cfg)
    assert np.array_equal(merged["hist"], want_hist) and merged["mIoU"] == oio.mean_iou(want_hist)


@pytest.mark.parametrize("factor", [2.0, 1.5])
def test_loader_resizes_frames_like_the_reference(factor):
    """Frames that do not arrive at config.SCALES go through `resize` (lib/utils/image.py:194-213: short side to the
    target, cv2.INTER_LINEAR) before `transform`, on the device: TestLoader's `data` must equal transform(cv2.resize(frame))
    bit for bit and im_info must carry the scale (get_image, image.py:40-44)."""
    cv2 = pytest.importorskip("cv2")
    cfg = loader.default_config(key_frame_interval=2, scales=(H, W))
    hs, ws = int(H * factor), int(W * factor)
    g = torch.Generator().manual_seed(9)
    frames = torch.randint(0, 256, (2, hs, ws, 3), generator=g, dtype=torch.uint8)
    roidb = [{"pattern": None, "frames": frames, "frame_seg_len": 2, "frame_id": 0}]
    data = loader.TestLoader(roidb, cfg, device="cuda:0")
    scale = float(H) / float(hs)
    seen = 0
    for t, (im_info, flag, batch) in enumerate(data):
        ref_im = cv2.resize(frames[t].numpy(), None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
        want = oio.transform(ref_im, synthetic.PIXEL_MEANS_BGR)
        got = batch.data[0][0]
        assert tuple(got.shape) == (1, 3, H, W) and np.array_equal(got.cpu().numpy(), want)
        assert abs(float(im_info[0][0, 2]) - scale) < 1e-7
        seen += 1
    assert seen == 2


def test_pred_eval_multiprocess_two_gpus(tmp_path):
    """One predictor pair + one TestLoader per GPU, driven from one process (tester.py:305-316): needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = loader.default_config(key_frame_interval=INTERVAL, scales=(H, W))
    roidb = _roidb()
    params = {k: v.numpy() for k, v in synthetic.make_params("dff").items()}
    arg, aux = params_io.split_arg_aux(params)
    sym = predictor.dff_deeplab()
    shapes = [[("data", (1, 3, H, W)), ("data_key", (1, 3, H, W))]]
    keys, curs, datas = [], [], []
    for g in range(2):
        datas.append(loader.TestLoader([roidb[g]], cfg, device="cuda:%d" % g))
        a, x = dict(arg), dict(aux)                     # distinct dicts -> distinct engines per device
        keys.append(predictor.Predictor(sym.get_key_test_symbol(cfg), datas[g].data_name, [], context=[predictor.gpu(g)],
                                        max_data_shapes=shapes, arg_params=a, aux_params=x))
        curs.append(predictor.Predictor(sym.get_cur_test_symbol(cfg), datas[g].data_name, [], context=[predictor.gpu(g)],
                                        max_data_shapes=shapes, arg_params=a, aux_params=x, engine=keys[g].engine))
    merged = loader.pred_eval_multiprocess(2, keys, curs, datas, None, cfg)
    # the same two videos on one GPU
    single = loader.pred_eval(0, keys[0], curs[0], loader.TestLoader(roidb, cfg, device="cuda:0"), None, cfg)
    assert np.array_equal(merged["hist"], single["hist"])
    assert sorted(merged["frame_ids"]) == sorted(single["frame_ids"])
