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
