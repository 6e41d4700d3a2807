"""Per-layer parity of the R101-DCN key net (-m gpu): the internal output of EVERY layer of the key plan
(accel_debug_fetch: split fp16 NHWC -> fp32 NCHW) against the oracle's activation of the same name -- the check that
located round 2's findings (a single deformable-conv output pixel flipping at the image border; 256-wide tiles truncating
the cross terms) and that would localise the next one to a layer instead of to "the score volume differs"."""
import pytest
import torch

from accel_b200 import synthetic
from accel_b200.engine import Engine
from oracle import nets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(128, 256), (256, 384)])
def test_every_layer_of_the_key_net_matches_the_oracle(H, W):
    params = synthetic.make_params("dff")
    frame = synthetic.make_frames(1, H, W, stream=4)[0]
    eng = Engine("dff", H, W, params=params)
    dev = eng.torch_device
    feat = torch.empty(eng.feat_shape, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        acts = nets.resnet_dcn_101_layers(params, frame)
        assert torch.equal(acts["res5c_branch2c"], nets.resnet_dcn_101(params, frame))      # same code path as the graph
    eng.key_forward(frame.to(dev), feat, None, label)
    worst = ("", 0.0)
    for name, ref in acts.items():
        got = eng.fetch_layer("key", name)
        assert tuple(got.shape) == tuple(ref.shape), name
        rel = (got.cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        assert rel < 3e-5, "%s: max-abs error %.3e of |ref|max" % (name, rel)
        if rel > worst[1]:
            worst = (name, rel)
    print("per-layer parity %dx%d: %d layers, worst %s at %.2e of |ref|max" % (H, W, len(acts), worst[0], worst[1]))
    with pytest.raises(RuntimeError):
        eng.fetch_layer("key", "no_such_layer")
    eng.close()
