"""GPU parity of the ingest / metric kernels (accel_preprocess, accel_confusion) and of the overlapped
host-to-host VideoPipeline against the plain per-frame loop.  Integer / byte work: bit-exact."""
import numpy as np
import pytest
import torch

from accel_b200 import scheduler, synthetic
from oracle import io as oio

pytestmark = pytest.mark.gpu
MEANS = synthetic.PIXEL_MEANS_BGR


@pytest.mark.parametrize("h,w", [(1, 1), (3, 5), (16, 20), (128, 256), (1024, 2048)])
def test_preprocess_bit_exact(h, w):
    from accel_b200 import engine as E
    g = torch.Generator().manual_seed(h * 131 + w)
    im = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
    out = E.preprocess(im.cuda())
    want = oio.transform(im.numpy(), MEANS)
    assert out.shape == (1, 3, h, w)
    assert np.array_equal(out.cpu().numpy(), want)


def test_preprocess_all_byte_values_and_custom_means():
    from accel_b200 import engine as E
    v = torch.arange(256, dtype=torch.uint8)
    im = torch.stack([v, v.flip(0), v.roll(7)], dim=-1).reshape(16, 16, 3).contiguous()
    means = (0.1, 200.7, 1e-3)
    out = E.preprocess(im.cuda(), pixel_means_bgr=means)
    assert np.array_equal(out.cpu().numpy(), oio.transform(im.numpy(), means))


def test_preprocess_rejects_bad_input():
    from accel_b200 import engine as E
    with pytest.raises(TypeError):
        E.preprocess(torch.zeros(4, 4, 3, dtype=torch.uint8))                 # host tensor
    with pytest.raises(TypeError):
        E.preprocess(torch.zeros(4, 4, 3, device="cuda"))                     # not uint8


@pytest.mark.parametrize("n,k", [(1, 19), (15, 19), (16, 19), (1000, 5), (1024 * 2048, 19), (1024 * 2048 + 13, 19)])
def test_confusion_bit_exact(n, k):
    from accel_b200 import engine as E
    g = torch.Generator().manual_seed(n + k)
    pred = torch.randint(0, k, (n,), generator=g, dtype=torch.uint8)
    label = torch.randint(0, k, (n,), generator=g, dtype=torch.uint8)
    label[torch.rand(n, generator=g) < 0.25] = 255
    hist = E.confusion(pred.cuda(), label.cuda(), None, k)
    want = oio.fast_hist(pred.numpy(), label.numpy(), k)
    assert np.array_equal(hist.cpu().numpy(), want)
    # accumulation (`hist += curr_hist`, demo.py:272)
    E.confusion(pred.cuda(), label.cuda(), hist, k)
    assert np.array_equal(hist.cpu().numpy(), 2 * want)


def test_confusion_unaligned_views_and_all_ignored():
    from accel_b200 import engine as E
    g = torch.Generator().manual_seed(5)
    base_p = torch.randint(0, 19, (4099,), generator=g, dtype=torch.uint8).cuda()
    base_l = torch.randint(0, 19, (4099,), generator=g, dtype=torch.uint8).cuda()
    p, l = base_p[3:], base_l[3:]                                              # 3-byte offset: scalar path
    hist = E.confusion(p.contiguous() if not p.is_contiguous() else p, l, None, 19)
    assert np.array_equal(hist.cpu().numpy(), oio.fast_hist(p.cpu().numpy(), l.cpu().numpy(), 19))
    ign = torch.full((4096,), 255, dtype=torch.uint8).cuda()
    assert int(E.confusion(base_p[:4096], ign, None, 19).sum()) == 0


@pytest.mark.parametrize("schedule", ["chained", "unchained"])
def test_video_pipeline_equals_plain_loop(schedule):
    """Host uint8 frames -> host labels through the overlapped pipeline == segment_frame on resident fp32
    frames, label for label; the accumulated confusion matrix equals fast_hist summed over frames."""
    from accel_b200 import engine as E
    from accel_b200.engine import Engine
    H, W, T, I = 128, 256, 7, 3
    eng = Engine("18", H, W, params=synthetic.make_params("18"))
    dev = eng.torch_device
    frames_u8 = synthetic.make_frames_u8(T, H, W, stream=2)
    host = [f.contiguous().pin_memory() for f in frames_u8]
    g = torch.Generator().manual_seed(9)
    gts = [torch.randint(0, 19, (H, W), generator=g, dtype=torch.uint8) for _ in range(T)]
    for t in gts:
        t[::5, ::3] = 255
    pipe = scheduler.VideoPipeline(eng, I, schedule)
    labels = pipe.segment_video(host, gt_labels=[t.to(dev) for t in gts])
    # plain loop
    state = scheduler.StreamState(eng)
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    want_hist = np.zeros((19, 19), dtype=np.int64)
    for i, f in enumerate(frames_u8):
        data = synthetic.transform(f).to(dev)
        scheduler.segment_frame(eng, state, data, I, schedule, lab)
        ref = lab.cpu().numpy()
        assert np.array_equal(labels[i].numpy(), ref), "frame %d" % i
        want_hist += oio.fast_hist(ref.flatten(), gts[i].numpy().flatten(), 19)
    assert np.array_equal(pipe.hist.cpu().numpy(), want_hist)
    # a second video through the same pipeline starts with a key frame again
    labels2 = pipe.segment_video(host[:2])
    assert np.array_equal(labels2[0].numpy(), labels[0].numpy())
    assert np.array_equal(labels2[1].numpy(), labels[1].numpy())
    eng.close()


@pytest.mark.parametrize("schedule", ["chained", "unchained"])
@pytest.mark.parametrize("interval", [1, 2, 4])
def test_key_lookahead_equals_sequential_loop(schedule, interval):
    """segment_frame with the next key frame handed in advance (key plan on a second stream) produces the same
    label maps and features as the strictly sequential loop."""
    from accel_b200.engine import Engine
    H, W, T = 128, 256, 9
    eng = Engine("dff", H, W, params=synthetic.make_params("dff"))
    dev = eng.torch_device
    frames = [f.to(dev) for f in synthetic.make_frames(T, H, W, stream=6)]
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    ref_labels, ref_feats = [], []
    st = scheduler.StreamState(eng)
    for f in frames:
        scheduler.segment_frame(eng, st, f, interval, schedule, lab)
        ref_labels.append(lab.cpu().numpy().copy())
        ref_feats.append(st.feat_in.cpu().numpy().copy())
    st = scheduler.StreamState(eng)
    for i, f in enumerate(frames):
        nk = frames[i + interval] if (i % interval == 0 and i + interval < T) else None
        scheduler.segment_frame(eng, st, f, interval, schedule, lab, next_key_data=nk)
        assert np.array_equal(lab.cpu().numpy(), ref_labels[i]), "frame %d" % i
        assert np.array_equal(st.feat_in.cpu().numpy(), ref_feats[i]), "frame %d" % i
    eng.close()


@pytest.mark.parametrize("version", ["dff", "101"])
def test_video_pipeline_whole_intervals(version):
    """VideoPipeline.submit_interval (host uint8 frames of a whole key interval in, host label maps out, two buffer sets in
    flight) == accel_interval_forward on resident fp32 frames, label for label, over several intervals."""
    from accel_b200.engine import Engine
    H, W, I, N = 128, 256, 3, 4
    eng = Engine(version, H, W, params=synthetic.make_params(version), interval=I)
    dev = eng.torch_device
    frames_u8 = synthetic.make_frames_u8(N * I, H, W, stream=6)
    host = [f.contiguous().pin_memory() for f in frames_u8]
    out = [torch.empty(H, W, dtype=torch.uint8).pin_memory() for _ in range(N * I)]
    pipe = scheduler.VideoPipeline(eng, I, "chained", batched=True)
    evs = []
    for k in range(N):
        evs += pipe.submit_interval(host[k * I:(k + 1) * I], out[k * I:(k + 1) * I])
    pipe.sync()
    assert all(e.query() for e in evs)
    st = scheduler.StreamState(eng)
    labels = torch.empty(I, H, W, dtype=torch.uint8, device=dev)
    for k in range(N):
        fr = [synthetic.transform(f).to(dev) for f in frames_u8[k * I:(k + 1) * I]]
        scheduler.segment_interval(eng, st, fr, labels)
        for t in range(I):
            assert np.array_equal(out[k * I + t].numpy(), labels[t].cpu().numpy()), "interval %d frame %d" % (k, t)
    with pytest.raises(ValueError):
        pipe.submit_interval(host[:2], out[:2])
    with pytest.raises(ValueError):
        scheduler.VideoPipeline(eng, I + 1, "chained", batched=True)
    eng.close()


# ----------------------------------------------------------------------------- cv2.resize ingest (lib/utils/image.py:194-222)
def _resize_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_vectors.npz"))
    return [(g["src_%d" % i], float(g["fx_%d" % i]), g["dst_%d" % i]) for i in range(int(g["n"]))]


def test_resize_matches_cv2_golden_vectors():
    """accel_resize_bgr against outputs of the real cv2.resize(INTER_LINEAR) (tests/golden/make_resize_vectors.py)."""
    from accel_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    for src, fx, dst in _resize_cases():
        s = torch.from_numpy(src).cuda().contiguous()
        dh, dw = C.c_int(), C.c_int()
        assert lib.accel_resize_size(src.shape[0], src.shape[1], fx, fx, C.byref(dh), C.byref(dw)) == 0
        assert (dh.value, dw.value) == dst.shape[:2], (src.shape, fx)
        out = torch.empty(dh.value, dw.value, 3, dtype=torch.uint8, device="cuda")
        assert lib.accel_resize_bgr(C.c_void_p(s.data_ptr()), src.shape[0], src.shape[1], fx, fx, C.c_void_p(out.data_ptr()), None) == 0
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), dst), (src.shape, fx)


def test_resize_full_size_against_live_cv2_and_reference_scale_rule():
    """Frame-sized inputs through engine.resize (the reference's scale rule, image.py:204-210) against cv2 itself where
    it is installed: a 2048x4096 frame to 1024x2048 (exact 2x: OpenCV's INTER_AREA path), 1080x1920 (short side to
    1024), 600x800 (long side capped)."""
    cv2 = pytest.importorskip("cv2")
    from accel_b200 import engine as E
    rng = np.random.default_rng(5)
    for (h, w), (target, max_size) in (((2048, 4096), (1024, 2048)), ((1080, 1920), (1024, 2048)), ((600, 800), (1024, 1280))):
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out, scale = E.resize(torch.from_numpy(im).cuda(), target, max_size)
        ref_scale = E.resize_scale(h, w, target, max_size)
        ref = cv2.resize(im, None, None, fx=ref_scale, fy=ref_scale, interpolation=cv2.INTER_LINEAR)
        assert scale == ref_scale and tuple(out.shape) == ref.shape
        assert np.array_equal(out.cpu().numpy(), ref), (h, w)
    out, _ = E.resize(torch.from_numpy(im).cuda(), 1024, 1280, stride=32)          # IMAGE_STRIDE padding, image.py:215-222
    assert out.shape[0] % 32 == 0 and out.shape[1] % 32 == 0 and int(out[ref.shape[0]:].sum()) == 0
