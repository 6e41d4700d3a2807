"""TestLoader-shaped video iterator and the segmentation `pred_eval`.

Mirrors dff_deeplab/core/loader.py:197-303 (`TestLoader`: which frame is a key frame, what `data_key` is,
the (im_info, key_frame_flag, DataBatch) triple it yields) and dff_deeplab/core/tester.py:203-316
(`pred_eval` / `pred_eval_multiprocess`: key predictor on flags 0/1, cur predictor with the cached key
feature on flag 2 -- the UN-chained schedule).  The reference's pred_eval body is still detection code
(NMS, all_boxes); the segmentation variant here keeps its control flow and replaces the detection
post-processing with what demo.py does with a score map (argmax -> uint8, fast_hist, demo.py:238-272).

Frames come from `roidb` entries -- one per video snippet, as in the reference:
    {'pattern': 'dir/%06d.png' or None, 'frames': (T,H,W,3) uint8 BGR array or list (used when no reader),
     'frame_seg_len': T, 'frame_id': first global frame id, 'labels': optional {frame offset: (H,W) uint8}}
`reader(path) -> (H,W,3) uint8 BGR` stands in for cv2.imread (no OpenCV in this image).
"""
from __future__ import annotations

import numpy as np
import torch

from .netspec import FEAT_DIM, NUM_CLASSES
from .predictor import DataBatch


class AttrDict(dict):
    """easydict.EasyDict stand-in for the `config` object the reference passes around."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_config(key_frame_interval=5, scales=(1024, 2048)):
    """The fields of experiments/dff_deeplab/cfgs/dff_deeplab_vid_demo.yaml this path reads."""
    return AttrDict(
        SCALES=[tuple(scales)],
        network=AttrDict(PIXEL_MEANS=np.array([103.06, 115.90, 123.15]), IMAGE_STRIDE=0, DFF_FEAT_DIM=FEAT_DIM),
        dataset=AttrDict(NUM_CLASSES=NUM_CLASSES),
        TEST=AttrDict(KEY_FRAME_INTERVAL=int(key_frame_interval)))


class TestLoader:
    """core/loader.py:197-303.  batch_size 1, no shuffle (pred_eval asserts it)."""
    __test__ = False                                   # not a pytest class

    def __init__(self, roidb, config, batch_size=1, shuffle=False, has_rpn=False, device="cuda:0", reader=None):
        if batch_size != 1:
            raise ValueError("TestLoader: batch_size must be 1")
        self.cfg, self.roidb, self.batch_size, self.shuffle, self.has_rpn = config, roidb, batch_size, shuffle, has_rpn
        self.device, self.reader = torch.device(device), reader
        self.size = int(np.sum([x["frame_seg_len"] for x in self.roidb]))
        self.index = np.arange(self.size)
        self.data_name = ["data", "im_info", "data_key", "feat_key"]
        self.label_name = None
        self.cur_roidb_index = 0
        self.cur_frameid = 0
        self.data_key = None
        self.key_frameid = 0
        self.cur_seg_len = 0
        self.key_frame_flag = -1
        self.cur = 0
        self.data = None
        self.label = []
        self.im_info = None
        self._feat_placeholder = torch.zeros(1, config.network.DFF_FEAT_DIM, 1, 1, device=self.device)
        self.reset()
        if self.size:
            self.get_batch()

    # ---- mx.io.DataIter surface ---------------------------------------------------------------------
    @property
    def provide_data(self):
        return [[(k, tuple(v.shape)) for k, v in zip(self.data_name, idata)] for idata in self.data]

    @property
    def provide_label(self):
        return [None for _ in range(len(self.data))]

    @property
    def provide_data_single(self):
        return [(k, tuple(v.shape)) for k, v in zip(self.data_name, self.data[0])]

    @property
    def provide_label_single(self):
        return None

    def reset(self):
        self.cur = 0
        self.cur_roidb_index = 0
        self.cur_frameid = 0
        self.key_frameid = 0
        if self.shuffle:
            np.random.shuffle(self.index)

    def iter_next(self):
        return self.cur < self.size

    def __iter__(self):
        return self

    def __next__(self):
        return self.next()

    def next(self):                                                   # loader.py:259-276
        if not self.iter_next():
            raise StopIteration
        self.get_batch()
        frame = (self.cur_roidb_index, self.cur_frameid)
        self.cur += self.batch_size
        self.cur_frameid += 1
        if self.cur_frameid == self.cur_seg_len:
            self.cur_roidb_index += 1
            self.cur_frameid = 0
            self.key_frameid = 0
        elif self.cur_frameid - self.key_frameid == self.cfg.TEST.KEY_FRAME_INTERVAL:
            self.key_frameid = self.cur_frameid
        batch = DataBatch(data=self.data, label=self.label, pad=self.getpad(), index=self.getindex(),
                          provide_data=self.provide_data, provide_label=self.provide_label)
        batch.frame = frame                                           # (video, offset): lets pred_eval find the label
        return self.im_info, self.key_frame_flag, batch

    def getindex(self):
        return self.cur // self.batch_size

    def getpad(self):
        return max(self.cur + self.batch_size - self.size, 0)

    # ---- frame ingest: get_rpn_testbatch -> get_image -> resize + transform (lib/utils/image.py) --------
    def _frame_u8(self, rec, frameid):
        if self.reader is not None and rec.get("pattern"):
            im = self.reader(rec["pattern"] % frameid)
        else:
            im = rec["frames"][frameid]
        im = torch.as_tensor(np.ascontiguousarray(im)) if not isinstance(im, torch.Tensor) else im
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("frames must be (H,W,3) uint8 BGR")
        return im

    def _ingest(self, rec, frameid):
        """get_rpn_testbatch -> get_image -> resize + transform (lib/utils/image.py:194-235) for one frame: the uint8
        BGR image goes to the device, frames that do not arrive at config.SCALES are resized there (accel_resize_bgr:
        cv2.INTER_LINEAR bit for bit) and `accel_preprocess` builds the fp32 `data` tensor (no CPU path)."""
        from . import engine as E
        im = self._frame_u8(rec, frameid).to(self.device, non_blocking=True).contiguous()
        target, max_size = self.cfg.SCALES[0]
        self.im_scale = 1.0
        if min(im.shape[0], im.shape[1]) != target or max(im.shape[0], im.shape[1]) > max_size:
            im, self.im_scale = E.resize(im, target, max_size, stride=int(getattr(self.cfg.network, "IMAGE_STRIDE", 0) or 0))
        return E.preprocess(im.contiguous(), None, tuple(float(m) for m in self.cfg.network.PIXEL_MEANS))

    def get_batch(self):                                              # loader.py:278-303
        rec = self.roidb[self.cur_roidb_index]
        self.cur_seg_len = rec["frame_seg_len"]
        data = self._ingest(rec, self.cur_frameid)
        im_info = torch.tensor([[data.shape[2], data.shape[3], getattr(self, "im_scale", 1.0)]], dtype=torch.float32)
        if self.key_frameid == self.cur_frameid:                      # key frame
            self.data_key = data.clone()
            self.key_frame_flag = 0 if self.key_frameid == 0 else 1
        else:
            self.key_frame_flag = 2
        self.data = [[data, im_info, self.data_key, self._feat_placeholder]]
        self.im_info = [im_info.numpy()]


def pred_eval(gpu_id, key_predictor, cur_predictor, test_data, imdb=None, cfg=None, vis=False, thresh=1e-4, logger=None,
              ignore_cache=True, keep_labels=False):
    """Segmentation variant of core/tester.py:203-303.  Returns {'hist': (n,n) int64 confusion counts over the
    frames that have a ground-truth label map, 'frame_ids': global id per processed frame, 'labels': list of
    uint8 label maps when keep_labels}.  The key feature is reused for the whole interval (tester.py:252-256)."""
    from . import engine as E
    from .predictor import im_segment
    assert vis or not test_data.shuffle
    n = cfg.dataset.NUM_CLASSES if cfg is not None else NUM_CLASSES
    dev = test_data.device
    num_images = test_data.size
    roidb_frame_ids = [x["frame_id"] for x in test_data.roidb]
    frame_ids = np.zeros(num_images, dtype=np.int64)
    hist = torch.zeros(n, n, dtype=torch.int64, device=dev)
    labels = []
    roidb_idx, roidb_offset, idx = -1, -1, 0
    feat = None
    for im_info, key_frame_flag, data_batch in test_data:
        if key_frame_flag != 2:
            output_all, feat = im_segment(key_predictor, data_batch)
        else:
            data_batch.data[0][-1] = feat
            data_batch.provide_data[0][-1] = ("feat_key", tuple(feat.shape))
            output_all, _ = im_segment(cur_predictor, data_batch)
        if key_frame_flag == 0:
            roidb_idx += 1
            roidb_offset = 0
        else:
            roidb_offset += 1
        frame_ids[idx] = roidb_frame_ids[roidb_idx] + roidb_offset
        pred = output_all[0]["label_output"]                           # uint8 argmax (demo.py:238,252)
        gt = (test_data.roidb[roidb_idx].get("labels") or {}).get(roidb_offset)
        if gt is not None:
            gt = torch.as_tensor(np.ascontiguousarray(gt), dtype=torch.uint8).to(dev)
            E.confusion(pred.contiguous(), gt.contiguous(), hist, n)   # hist += fast_hist(...)  (demo.py:270-272)
        if keep_labels:
            labels.append(pred.cpu().numpy().copy())
        idx += test_data.batch_size
        if logger:
            logger.info("testing {}/{}".format(idx, num_images))
    return {"gpu_id": gpu_id, "hist": hist.cpu().numpy(), "frame_ids": frame_ids, "labels": labels}


def merge_results(results):
    """Host-side merge of the per-GPU results (lib/dataset/imagenet_vid.py:199 merges detections the same way)."""
    hist = sum(r["hist"] for r in results)
    frame_ids = np.concatenate([r["frame_ids"] for r in results]) if results else np.zeros(0, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.true_divide(np.diag(hist), (hist.sum(1) + hist.sum(0) - np.diag(hist)))       # demo.py:55-56
    return {"hist": hist, "frame_ids": frame_ids, "ious": iu * 100,
            "mIoU": round(float(np.nanmean(iu)) * 100, 2) if np.isfinite(iu).any() else float("nan")}


def pred_eval_multiprocess(gpu_num, key_predictors, cur_predictors, test_datas, imdb=None, cfg=None, vis=False, thresh=1e-4,
                           logger=None, ignore_cache=True):
    """tester.py:305-316.  One predictor pair + one TestLoader per GPU; the loops are issued from one host
    thread per GPU (CUDA work is asynchronous; the reference uses a process pool because MXNet's Python
    front end blocks)."""
    if gpu_num == 1:
        res = [pred_eval(0, key_predictors[0], cur_predictors[0], test_datas[0], imdb, cfg, vis, thresh, logger, ignore_cache)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=gpu_num) as pool:
            futs = [pool.submit(pred_eval, i, key_predictors[i], cur_predictors[i], test_datas[i], imdb, cfg, vis, thresh,
                                logger, ignore_cache) for i in range(gpu_num)]
            res = [f.result() for f in futs]
    out = merge_results(res)
    if logger:
        logger.info("evaluate segmentation: mIoU {:.3f}".format(out["mIoU"]))
    return out
