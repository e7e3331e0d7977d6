"""CPU, world_size 2 over gloo: the sharding and the one collective of the pipeline.
No kernels run here -- this covers the host logic of the N>1 path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_super_resolution_b200.pipeline import gather_frames, gather_frames_ragged, shard_windows


def test_shard_windows_covers_all_once():
    for n, g in [(298, 8), (8, 8), (7, 4), (1, 2), (37, 3)]:
        seen = []
        for r in range(g):
            seen += list(shard_windows(n, g, r))
        assert seen == list(range(n)), (n, g)
    # contiguous chunks: the recurrence of window k+1 on window k stays on one rank
    assert list(shard_windows(298, 8, 0)) == list(range(0, 38))
    assert list(shard_windows(298, 8, 7)) == list(range(266, 298))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, H, W = 3, 8, 12
        idx = list(shard_windows(n * world, world, rank))
        local = torch.stack([torch.full((H, W, 3), i, dtype=torch.uint8) for i in idx])
        allf = gather_frames(local)
        ok = tuple(allf.shape) == (n * world, H, W, 3) and all(int(allf[i, 0, 0, 0]) == i for i in range(n * world))
        # shards of unequal length (C5: 38/38/.../32 windows): counts exchanged, padded gather, padding dropped
        mine = list(shard_windows(5, world, rank))               # 3 + 2 windows
        local = torch.stack([torch.full((H, W, 3), 10 + i, dtype=torch.uint8) for i in mine])
        rag = gather_frames_ragged(local)
        ok = ok and tuple(rag.shape) == (5, H, W, 3) and all(int(rag[i, 0, 0, 0]) == 10 + i for i in range(5))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gather_frames_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_gather_frames_single_process_is_identity():
    x = torch.zeros((2, 4, 4, 3), dtype=torch.uint8)
    assert gather_frames(x) is x
