"""World-size-2 test of the multi-GPU host logic on CPU (gloo): contiguous shard ranges cover the input exactly once,
the index-image broadcast protocol delivers identical bytes, and the ordered merge reproduces the single-process result."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_reads, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import _libs as L
    from airlift_b200 import dist as D
    # 1. "index image": rank 0 owns the bytes, everybody must end up with the same digest
    shape = D.broadcast_object((11, 21, 0, 3, 12345) if rank == 0 else None)
    assert shape == (11, 21, 0, 3, 12345)
    rng = np.random.default_rng(5)
    for nbytes in (1, 4096, 1 << 20):
        want = torch.from_numpy(rng.integers(0, 256, nbytes, dtype=np.uint8))
        buf = want.clone() if rank == 0 else torch.zeros(nbytes, dtype=torch.uint8)
        dist.broadcast(buf, src=0)
        assert torch.equal(buf, want)
    # 2. shard the reads, "map" them (oracle sketch digest per read), merge in order on rank 0
    rng = np.random.default_rng(9)
    reads = [L.rand_seq(rng, int(rng.integers(30, 200))) for _ in range(n_reads)]
    b, e = D.shard_range(n_reads, rank, world)
    local = [int(L.orc_sketch(r, 11, 21)["x"].sum() & np.uint64(0xffffffff)) for r in reads[b:e]]
    merged = D.gather_in_order(local, dst=0)
    if rank == 0:
        single = [int(L.orc_sketch(r, 11, 21)["x"].sum() & np.uint64(0xffffffff)) for r in reads]
        ret["ok"] = merged == single
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from airlift_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 101, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
