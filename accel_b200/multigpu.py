"""Multi-GPU plumbing for the per-frame path: one process per GPU, one independent video stream (or a
greedy shard of whole streams) per process, and NO data-path collective -- the reference runs one
predictor pair per GPU and merges results on the host (dff_rfcn/function/test_rcnn.py:60-82,
dff_deeplab/core/tester.py:305-314).  The only exchange is the end-of-run metric reduction below, over
`torch.distributed` (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    """Joins the torchrun rendezvous (MASTER_ADDR/MASTER_PORT from the environment).  No-op for world size 1."""
    rank, size, _ = world()
    if size > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    return rank, size


def barrier():
    if dist.is_initialized():
        dist.barrier()


def gather_rows(values, device="cpu"):
    """all_gather of one small float64 row per rank -> (world, len(values)) tensor on every rank."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if not dist.is_initialized():
        return t.unsqueeze(0).cpu()
    rows = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(rows, t)
    return torch.stack(rows).cpu()


def aggregate_throughput(frames_per_rank, ms_per_rank):
    """Whole-job frames/s: every stream's frames over the slowest rank's device time."""
    slowest = max(float(m) for m in ms_per_rank)
    return sum(float(f) for f in frames_per_rank) / (slowest / 1000.0), slowest


def reduce_confusion(hist):
    """Sum of the per-rank confusion matrices (fast_hist, dff_deeplab/demo.py:50-53) on every rank."""
    if dist.is_initialized():
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def per_class_iu(hist):
    """dff_deeplab/demo.py:55-56."""
    hist = hist.double()
    return torch.diag(hist) / (hist.sum(1) + hist.sum(0) - torch.diag(hist))
