"""Keyframe scheduling for the Accel hot path: the control flow of the reference's two loops.

chained   -- dff_deeplab/demo.py:165-250.  Frame idx is a key frame when idx % interval == 0;
             otherwise the cur graph runs with data_key = the PREVIOUS frame and feat_key = the
             feature returned by the previous call (key feature or warped feature).
unchained -- dff_deeplab/core/loader.py:259-303 + core/tester.py:246-256 (pred_eval).  data_key is
             the KEY frame and feat_key stays the key frame's res5c feature for the whole interval.

`key_frame_flags` reproduces TestLoader's flag stream (0 first key frame of a video, 1 later key
frames, 2 non-key frames).  `shard_streams` is the per-GPU assignment of whole videos
(dff_rfcn/function/test_rcnn.py:60-67: greedy, least-loaded GPU first).
"""
from __future__ import annotations

import numpy as np
import torch

from .predictor import DataBatch, feat_key_placeholder, im_segment

SCHEDULES = ("chained", "unchained")


def key_frame_flags(num_frames, interval):
    if interval < 1:
        raise ValueError("Invalid interval %d - must be >=1" % interval)     # demo.py:120-121
    flags, key = [], 0
    for cur in range(num_frames):
        if cur - key == interval:                                             # loader.py:268-269
            key = cur
        flags.append((0 if key == 0 else 1) if key == cur else 2)
    return flags


def shard_streams(stream_lengths, num_gpus):
    """Greedy assignment of whole streams to the least-loaded GPU (test_rcnn.py:60-67).
    Returns a list (per GPU) of stream indices."""
    load = [0] * num_gpus
    shards = [[] for _ in range(num_gpus)]
    for idx, n in enumerate(stream_lengths):
        g = int(np.argmin(load))
        shards[g].append(idx)
        load[g] += int(n)
    return shards


def run_reference_loop(key_predictor, cur_predictor, frames, interval, version, schedule="chained"):
    """The demo.py:228-250 loop, literally, over Predictor objects: returns per-frame dicts with the
    score volume and the argmax'ed uint8 label map (np.uint8(argmax), demo.py:245,252)."""
    if schedule not in SCHEDULES:
        raise ValueError("schedule must be one of %s" % (SCHEDULES,))
    dev = frames[0].device
    results = []
    feat = None
    prev = key_img = None
    for idx, data in enumerate(frames):
        if prev is None:
            prev = data
        is_key = idx % interval == 0
        if is_key:
            key_img = data
        data_key = prev if schedule == "chained" else key_img
        batch = DataBatch(data=[[data, data_key, feat_key_placeholder(dev)]])
        if is_key:
            output_all, feat_new = im_segment(key_predictor, batch)
            score = output_all[0]["croped_score_output"]
            feat = feat_new
        else:
            batch.data[0][-1] = feat                                           # demo.py:241
            output_all, feat_new = im_segment(cur_predictor, batch)
            output_key = "croped_score_output" if version in ("101", "dff") else "correction_output"
            score = output_all[0][output_key]
            if schedule == "chained":
                feat = feat_new
        label = torch.argmax(score, dim=1).to(torch.uint8)[0]
        results.append({"is_key": is_key, "label": label.clone(), "score": score.clone(),
                        "label_output": output_all[0]["label_output"].clone(), "feat": feat.clone()})
        prev = data
    return results


class StreamState:
    """Per-video state the schedule carries between frames: {feat, data_key}, plus the
    buffers of the optional key-frame lookahead (see segment_frame)."""

    def __init__(self, engine, linear_head=False):
        """linear_head: carry G = W_fc6 * F (1024 channels) from frame to frame instead of the 2048-channel feature
        (the commuted L head, accel_cur_forward_lin).  Off by default: the graphs then are the reference's as
        written; bench.py and production pipelines turn it on."""
        dev = engine.torch_device
        self.linear_head = bool(linear_head) and engine.supports_linear_head
        shape = engine.g_shape if self.linear_head else engine.feat_shape
        self.carry_shape = shape
        self.feat = [torch.empty(shape, device=dev) for _ in range(2)]
        self.cur = 0
        self.key_frame = None
        self.prev_frame = None
        self.index = 0
        self.feat_in = self.feat[0]          # the feature the next cur frame reads
        # lookahead (allocated on first use)
        self.key_stream = None
        self.key_feat = None                 # two key-feature buffers, alternating per interval
        self.key_label = None
        self.key_slot = 0
        self.pending = None                  # {"data": tensor, "slot": int, "event": cuda event}


def _launch_lookahead(engine, state, next_key_data):
    """Starts the key plan of the NEXT interval's key frame on the state's key stream, after everything issued so
    far on the current stream (so the buffers it overwrites are no longer read)."""
    dev = engine.torch_device
    if state.key_stream is None:
        state.key_stream = torch.cuda.Stream(dev)
        state.key_feat = [torch.empty(state.carry_shape, device=dev) for _ in range(2)]
        state.key_label = [torch.empty(engine.height, engine.width, dtype=torch.uint8, device=dev) for _ in range(2)]
    slot = state.key_slot
    state.key_slot ^= 1
    main = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main)
    state.key_stream.wait_event(ready)
    with torch.cuda.stream(state.key_stream):
        if state.linear_head:
            engine.key_forward(next_key_data, None, None, state.key_label[slot], g_out=state.key_feat[slot])
        else:
            engine.key_forward(next_key_data, state.key_feat[slot], None, state.key_label[slot])
        done = torch.cuda.Event()
        done.record(state.key_stream)
    state.pending = {"data": next_key_data, "slot": slot, "event": done}


def segment_frame(engine, state, data, interval, schedule, label_out, score_out=None, next_key_data=None):
    """Production step (no score volume unless asked): one frame of one stream through the key/cur plans with
    double-buffered features.  Returns True when it was a key frame.

    Key-frame lookahead: a key frame depends on no earlier frame, so when the caller already holds the NEXT
    interval's key frame (a video file, a decode queue) it passes it as `next_key_data` with the current key
    frame; its key plan then runs on a second CUDA stream underneath this interval's cur frames (which are many
    small kernels) and is merely waited for when its turn comes.  Same plans, same inputs, same outputs as the
    sequential loop of demo.py:228-250 -- only the issue order on the GPU changes."""
    is_key = state.index % interval == 0
    dev = engine.torch_device
    if is_key:
        p = state.pending
        if p is not None and p["data"] is data and score_out is None:
            main = torch.cuda.current_stream(dev)
            main.wait_event(p["event"])
            label_out.copy_(state.key_label[p["slot"]], non_blocking=True)
            state.feat_in = state.key_feat[p["slot"]]
        else:
            if p is not None:                                   # a stale lookahead: let it drain before reusing buffers
                torch.cuda.current_stream(dev).wait_event(p["event"])
            if state.linear_head:
                engine.key_forward(data, None, score_out, label_out, g_out=state.feat[state.cur])
            else:
                engine.key_forward(data, state.feat[state.cur], score_out, label_out)
            state.feat_in = state.feat[state.cur]
        state.pending = None
        state.key_frame = data
        if next_key_data is not None:
            _launch_lookahead(engine, state, next_key_data)
    elif schedule == "chained":
        nxt = state.cur ^ 1
        fwd = engine.cur_forward_lin if state.linear_head else engine.cur_forward
        fwd(data, state.prev_frame, state.feat_in, state.feat[nxt], score_out, label_out)
        state.cur = nxt
        state.feat_in = state.feat[nxt]
    else:
        fwd = engine.cur_forward_lin if state.linear_head else engine.cur_forward
        fwd(data, state.key_frame, state.feat_in, None, score_out, label_out)
    state.prev_frame = data
    state.index += 1
    return is_key


def segment_interval(engine, state, frames, labels, scores=None):
    """One whole key interval of the chained schedule through the whole-interval plan (accel_interval_forward): the
    caller holds all `interval` frames (a video file, a decode queue; demo.py:165-185 preloads the clip).  Same
    graphs and inputs as `interval` calls of segment_frame; only the issue order on the GPU changes: the per-frame
    chains run concurrently.  labels: (interval,H,W) uint8 tensor or list of (H,W) tensors."""
    if state.pending is not None:                               # a lookahead left over from the frame-by-frame loop
        torch.cuda.current_stream(engine.torch_device).wait_event(state.pending["event"])
        state.pending = None
    engine.interval_forward(list(frames), [labels[t] for t in range(len(frames))], scores)
    state.key_frame = frames[0]
    state.prev_frame = frames[-1]
    state.index = 0                                             # the next frame is a key frame again


class VideoPipeline:
    """One video stream from HOST frames to HOST label maps -- the loop body of dff_deeplab/demo.py:228-252
    with the frame ingest of demo.py:170-175 moved onto the GPU.

    Per frame: the decoded uint8 BGR image (what cv2.imread + resize hand to `transform`, 3 bytes/pixel) is
    copied from pinned host memory on a copy stream, `accel_preprocess` turns it into the fp32 `data` tensor,
    the key or cur plan runs, and the uint8 label map is copied back on a third stream.  Copies of frame
    i+1 / labels of frame i-1 overlap the graphs of frame i; nothing else crosses PCIe.  Optionally the
    confusion matrix against ground-truth label maps is accumulated on the device (demo.py:270-272)."""

    def __init__(self, engine, interval, schedule="chained", pixel_means_bgr=None, depth=2, lookahead=True,
                 linear_head=False, batched=False):
        """batched: submit_interval() runs whole intervals through the engine's whole-interval plan."""
        if schedule not in SCHEDULES:
            raise ValueError("schedule must be one of %s" % (SCHEDULES,))
        if interval < 1:
            raise ValueError("Invalid interval %d - must be >=1" % interval)
        from . import engine as _E
        self._E = _E
        self.engine, self.interval, self.schedule, self.means = engine, int(interval), schedule, pixel_means_bgr
        self.lookahead = bool(lookahead)
        dev = engine.torch_device
        H, W = engine.height, engine.width
        self.dev, self.depth = dev, int(depth)
        self.state = StreamState(engine, linear_head=linear_head)
        self.nu8 = depth + 1                                                      # a key turn uploads two frames
        self.u8 = [torch.empty(H, W, 3, dtype=torch.uint8, device=dev) for _ in range(self.nu8)]
        self.f32 = [torch.empty(1, 3, H, W, device=dev) for _ in range(3)]       # cur / prev / key never collide
        self.keybuf = [torch.empty(1, 3, H, W, device=dev) for _ in range(2)] if self.lookahead else []
        self.label = [torch.empty(H, W, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.copy_in = torch.cuda.Stream(dev)
        self.copy_out = torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(self.nu8)]    # frame landed in u8[b]
        self.ev_free = [torch.cuda.Event() for _ in range(self.nu8)]  # u8[b] consumed by preprocess
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]     # label[b] written by the graph
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # label[b] copied to the host
        self.n = 0                                                    # frames submitted (label ring position)
        self.nu = 0                                                   # uploads issued (u8 ring position)
        self.kslot = 0
        self.pre_key = None                                           # {"host": pinned frame, "data": fp32 tensor}
        self.hist = torch.zeros(engine.num_classes, engine.num_classes, dtype=torch.int64, device=dev)
        self.batched = bool(batched)
        if self.batched:
            if schedule != "chained" or engine.interval != self.interval:
                raise ValueError("batched pipeline needs an Engine built with interval=%d and the chained schedule" % self.interval)
            I = self.interval
            self.iv_u8 = [[torch.empty(H, W, 3, dtype=torch.uint8, device=dev) for _ in range(I)] for _ in range(2)]
            self.iv_f32 = [[torch.empty(1, 3, H, W, device=dev) for _ in range(I)] for _ in range(2)]
            self.iv_label = [torch.empty(I, H, W, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.iv_in = [torch.cuda.Event() for _ in range(2)]      # set s: frames landed
            self.iv_free = [torch.cuda.Event() for _ in range(2)]    # set s: u8 frames consumed by preprocess
            self.iv_done = [torch.cuda.Event() for _ in range(2)]    # set s: label maps written
            self.iv_out = [torch.cuda.Event() for _ in range(2)]     # set s: label maps copied to the host
            self.iv_n = 0

    def submit_interval(self, frames_u8_host, labels_host):
        """Queues one whole key interval: `interval` pinned (H,W,3) uint8 BGR frames in, `interval` pinned (H,W) uint8
        label maps out (valid once the returned events have completed).  H2D copies of interval k+1 and D2H copies of
        interval k-1 overlap the plan of interval k (two buffer sets)."""
        I = self.interval
        if not self.batched or len(frames_u8_host) != I or len(labels_host) != I:
            raise ValueError("submit_interval needs a batched pipeline and exactly %d frames" % I)
        s = self.iv_n & 1
        main = torch.cuda.current_stream(self.dev)
        if self.iv_n >= 2:
            self.copy_in.wait_event(self.iv_free[s])
        with torch.cuda.stream(self.copy_in):
            for t in range(I):
                self.iv_u8[s][t].copy_(frames_u8_host[t], non_blocking=True)
            self.iv_in[s].record(self.copy_in)
        main.wait_event(self.iv_in[s])
        for t in range(I):
            self._E.preprocess(self.iv_u8[s][t], self.iv_f32[s][t], self.means)
        self.iv_free[s].record(main)
        if self.iv_n >= 2:
            main.wait_event(self.iv_out[s])                                  # label set s has left for the host
        segment_interval(self.engine, self.state, self.iv_f32[s], self.iv_label[s])
        self.iv_done[s].record(main)
        self.copy_out.wait_event(self.iv_done[s])
        with torch.cuda.stream(self.copy_out):
            for t in range(I):
                labels_host[t].copy_(self.iv_label[s][t], non_blocking=True)
            self.iv_out[s].record(self.copy_out)
        self.iv_n += 1
        return [self.iv_out[s]] * I

    def reset(self):
        """Start of a new video: the next frame is a key frame."""
        self.state.index = 0
        self.state.key_frame = self.state.prev_frame = None
        self.pre_key = None

    def _pick_f32(self):
        busy = {t.data_ptr() for t in (self.state.prev_frame, self.state.key_frame) if t is not None}
        for t in self.f32:
            if t.data_ptr() not in busy:
                return t
        raise RuntimeError("no free frame buffer")

    def _upload(self, frame_u8_host, into):
        """H2D of one uint8 frame on the copy stream + accel_preprocess on the current stream -> `into`."""
        b = self.nu % self.nu8
        main = torch.cuda.current_stream(self.dev)
        if self.nu >= self.nu8:
            self.copy_in.wait_event(self.ev_free[b])
        with torch.cuda.stream(self.copy_in):
            self.u8[b].copy_(frame_u8_host, non_blocking=True)
            self.ev_in[b].record(self.copy_in)
        main.wait_event(self.ev_in[b])
        self._E.preprocess(self.u8[b], into, self.means)
        self.ev_free[b].record(main)
        self.nu += 1
        return into

    def submit(self, frame_u8_host, label_host, gt_label=None, next_key_host=None):
        """Queues one frame.  frame_u8_host: pinned (H,W,3) uint8 BGR; label_host: pinned (H,W) uint8, valid
        after `sync()` (or once the returned event has completed); gt_label: optional CUDA uint8 (H,W)
        ground truth to accumulate `hist` against; next_key_host: with a key frame, the NEXT interval's key
        frame if the caller already has it (enables the key-frame lookahead of segment_frame).
        Returns (is_key, event)."""
        b = self.n % self.depth
        main = torch.cuda.current_stream(self.dev)
        key_turn = self.state.index % self.interval == 0
        if key_turn and self.pre_key is not None and self.pre_key["host"] is frame_u8_host:
            data = self.pre_key["data"]                                      # uploaded + preprocessed an interval ago
        else:
            data = self._upload(frame_u8_host, self._pick_f32())
        self.pre_key = None if key_turn else self.pre_key
        nk = None
        if key_turn and self.lookahead and next_key_host is not None:
            nk = self._upload(next_key_host, self.keybuf[self.kslot])
            self.kslot ^= 1
            self.pre_key = {"host": next_key_host, "data": nk}
        if self.n >= self.depth:
            main.wait_event(self.ev_out[b])                                  # label[b] has left for the host
        is_key = segment_frame(self.engine, self.state, data, self.interval, self.schedule, self.label[b],
                               next_key_data=nk)
        if gt_label is not None:
            self._E.confusion(self.label[b], gt_label, self.hist, self.engine.num_classes)
        self.ev_done[b].record(main)
        self.copy_out.wait_event(self.ev_done[b])
        with torch.cuda.stream(self.copy_out):
            label_host.copy_(self.label[b], non_blocking=True)
            self.ev_out[b].record(self.copy_out)
        self.n += 1
        return is_key, self.ev_out[b]

    def sync(self):
        self.copy_out.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
        if self.state.key_stream is not None:
            self.state.key_stream.synchronize()

    def segment_video(self, frames_u8_host, labels_host=None, gt_labels=None):
        """Whole clip: list of pinned uint8 frames -> list of pinned uint8 label maps (synchronised)."""
        self.reset()
        H, W = self.engine.height, self.engine.width
        T, I = len(frames_u8_host), self.interval
        if labels_host is None:
            labels_host = [torch.empty(H, W, dtype=torch.uint8).pin_memory() for _ in frames_u8_host]
        for i, f in enumerate(frames_u8_host):
            nxt = frames_u8_host[i + I] if (self.lookahead and i % I == 0 and i + I < T) else None
            self.submit(f, labels_host[i], None if gt_labels is None else gt_labels[i], next_key_host=nxt)
        self.sync()
        return labels_host


def confusion_matrix(pred, label, n):
    """fast_hist of demo.py:50-53 on the GPU (int64 n x n; rows = label)."""
    pred = pred.reshape(-1).long()
    label = label.reshape(-1).long()
    k = (label >= 0) & (label < n)
    return torch.bincount(n * label[k] + pred[k], minlength=n * n).reshape(n, n)
