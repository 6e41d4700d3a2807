"""World-size-2 gloo run (CPU) of the multi-GPU host logic: stream sharding, the single end-of-run metric
gather, max-over-ranks throughput and the confusion-matrix reduction.  No kernel runs here."""
import os
import socket

import torch
import torch.multiprocessing as mp

from accel_b200 import multigpu, scheduler


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, size, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, n = multigpu.init("gloo")
    assert (r, n) == (rank, size)
    lengths = [30, 10, 20, 25, 5]                               # frames per video stream
    shards = scheduler.shard_streams(lengths, size)
    mine = shards[rank]
    frames = sum(lengths[i] for i in mine)
    ms = 100.0 * frames * (1.0 + 0.5 * rank)                    # rank 1 is the slow one
    multigpu.barrier()
    rows = multigpu.gather_rows([frames, ms])
    fps, slowest = multigpu.aggregate_throughput(rows[:, 0].tolist(), rows[:, 1].tolist())
    # each rank labels its own pixels; the merged confusion matrix is the sum
    g = torch.Generator().manual_seed(rank)
    pred = torch.randint(0, 19, (64, 64), generator=g)
    label = torch.randint(0, 19, (64, 64), generator=g)
    label[0, :8] = 255                                          # ignore label, demo.py:51
    hist = multigpu.reduce_confusion(scheduler.confusion_matrix(pred, label, 19))
    if rank == 0:
        torch.save({"shards": shards, "rows": rows, "fps": fps, "slowest": slowest, "hist": hist}, out)
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_metric_reduction(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    shards = res["shards"]
    assert sorted(i for s in shards for i in s) == [0, 1, 2, 3, 4]          # every stream on exactly one GPU
    assert shards == [[0, 3], [1, 2, 4]]                                    # greedy least-loaded (test_rcnn.py:60-67)
    rows = res["rows"]
    assert rows.shape == (2, 2) and rows[:, 0].tolist() == [55.0, 35.0]
    assert res["slowest"] == max(rows[:, 1].tolist())
    assert abs(res["fps"] - 90.0 / (res["slowest"] / 1000.0)) < 1e-9
    hist = res["hist"]
    assert hist.shape == (19, 19) and int(hist.sum()) == 2 * (64 * 64 - 8)
    iu = multigpu.per_class_iu(hist)
    assert iu.shape == (19,) and bool(((iu >= 0) & (iu <= 1)).all())


def test_single_process_paths_need_no_group():
    rows = multigpu.gather_rows([5, 20.0])
    assert rows.shape == (1, 2)
    fps, slowest = multigpu.aggregate_throughput([5], [20.0])
    assert fps == 250.0 and slowest == 20.0
