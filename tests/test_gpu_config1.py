"""BASELINE config 1 on the CUDA path (-m gpu): DeepLab ResNet-18 (the correction network of Accel-18 + its head,
accel_rbranch_forward) on one 512x1024 frame through the Predictor shim, against the CPU oracle."""
import pytest
import torch

from accel_b200 import predictor as P
from accel_b200 import synthetic
from oracle import nets
from parity_util import label_report

pytestmark = pytest.mark.gpu
H, W = 512, 1024


@pytest.mark.parametrize("version", ["18", "50"])
def test_deeplab_single_frame_through_predictor_shim(version):
    params = synthetic.make_params(version)
    sym = P.deeplab(version).get_symbol(None, is_train=False)
    assert sym.list_outputs() == ["softmax_output"]
    pred = P.Predictor(sym, ["data"], [], context=[P.gpu(0)], provide_data=[[("data", (1, 3, H, W))]],
                       arg_params=dict(params), aux_params={})
    dev = pred.engine.torch_device
    data = synthetic.transform(synthetic.make_frames_u8(1, H, W, stream=3)[0])
    out = pred.predict(P.DataBatch(data=[[data.to(dev)]]))[0]
    with torch.no_grad():
        ref = nets.rbranch_forward(params, version, data)["croped_score_output"]
    rep = label_report(out["label_output"], out["croped_score_output"].cpu(), ref)
    print("config 1 (DeepLab-%s 512x1024): %r" % (version, rep))
    # deeplab/core/tester.py:85: argmax of the softmax output == the uint8 label map, off exact ties
    lab_sm = out["softmax_output"].argmax(dim=1)[0].to(torch.uint8)
    assert (lab_sm != out["label_output"]).float().mean().item() < 1e-4
    pred.engine.close()
