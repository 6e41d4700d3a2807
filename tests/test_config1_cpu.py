"""BASELINE config 1 (CPU plumbing, no GPU): the plain DeepLab ResNet-18 of deeplab/test.py on ONE synthetic 512x1024
frame -- here the correction network of Accel-18 with its own head (accel_18.py:199-227), which is that net -- run on
the CPU oracle through a Predictor-shaped object and the reference's per-image step
(deeplab/core/tester.py:84-85: predict -> argmax of `softmax_output`).  deeplab/test.py itself cannot run here
(MXNet, python2); the -m gpu twin of this test (tests/test_gpu_config1.py) runs the CUDA path against it."""
import numpy as np
import torch

from accel_b200 import synthetic
from accel_b200.predictor import DataBatch
from oracle import nets, ops

H, W = 512, 1024


class OraclePredictor:
    """deeplab/core/tester.py:22-35 Predictor, backed by the CPU oracle (test infrastructure)."""

    def __init__(self, version, data_names, arg_params, aux_params):
        assert list(data_names) == ["data"]                       # deeplab/function/test_deeplab.py:66
        self.version = version
        self.p = dict(arg_params)
        self.p.update(aux_params)

    def predict(self, data_batch):
        with torch.no_grad():
            out = nets.rbranch_forward(self.p, self.version, data_batch.data[0][0])
        return [{"softmax_output": torch.softmax(out["croped_score_output"], dim=1),
                 "croped_score_output": out["croped_score_output"]}]


def test_deeplab18_single_frame_plumbing_on_cpu():
    params = synthetic.make_params("18")
    arg = {k: v for k, v in params.items() if not k.endswith(("_moving_mean", "_moving_var"))}
    aux = {k: v for k, v in params.items() if k.endswith(("_moving_mean", "_moving_var"))}
    frame_u8 = synthetic.make_frames_u8(1, H, W, stream=3)[0]      # what cv2.imread would hand over: (H,W,3) BGR uint8
    data = synthetic.transform(frame_u8)                           # lib/utils/image.py:224-235
    assert data.shape == (1, 3, H, W) and data.dtype == torch.float32
    pred = OraclePredictor("18", ["data"], arg, aux)
    output_all = pred.predict(DataBatch(data=[[data]]))
    # deeplab/core/tester.py:85
    labels = [ops.argmax_channel(o["softmax_output"]) for o in output_all]
    assert labels[0].shape == (1, H, W)
    lab = np.uint8(labels[0][0])
    assert lab.max() < 19 and len(np.unique(lab)) > 1
    # argmax of the softmax == argmax of the score volume (softmax is monotonic per pixel), away from exact ties
    score = output_all[0]["croped_score_output"]
    top2 = score.topk(2, dim=1).values
    decided = ((top2[:, 0] - top2[:, 1])[0] > 1e-4).numpy()
    assert np.array_equal(lab[decided], np.uint8(ops.argmax_channel(score)[0])[decided])
    sm = output_all[0]["softmax_output"]
    assert torch.allclose(sm.sum(dim=1), torch.ones(1, H, W), atol=1e-5)
