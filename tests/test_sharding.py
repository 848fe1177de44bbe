"""CPU: the N>1 plumbing (window sharding + final consensus gather) with gloo, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from racon_gpu_b200.shard import ConsensusGather, assemble, gather_consensus, shard_range


def test_shard_ranges_partition_the_windows():
    for n in (0, 1, 7, 10000, 1000001):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_windows, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_windows, rank, world)
    stride = 64
    rng = np.random.default_rng(1234)  # same stream on every rank: row w is a function of w only
    all_len = rng.integers(1, stride, size=n_windows).astype(np.int32)
    all_rows = rng.integers(65, 90, size=(n_windows, stride)).astype(np.uint8)
    want = b"".join(all_rows[w, :all_len[w]].tobytes() for w in range(n_windows))
    flat, off = gather_consensus(all_rows[lo:hi].copy(), all_len[lo:hi].copy(), torch.device("cpu"))
    ok = True
    if rank == 0:
        ok = flat.tobytes() == want and (np.diff(off) == all_len).all()
    else:
        assert flat is None and off is None
    # the pipelined form: two steps in flight one after the other, only compact bytes travel
    g = ConsensusGather(torch.device("cpu"), hi - lo)
    for step in range(2):
        rows = np.roll(all_rows, step, axis=1)
        g.start(rows[lo:hi].copy(), all_len[lo:hi].copy())
        parts = g.wait()
        if rank == 0:
            f2, o2 = assemble(parts)
            ok = ok and f2.tobytes() == b"".join(rows[w, :all_len[w]].tobytes() for w in range(n_windows))
            ok = ok and sum(p[1].shape[0] for p in parts) == int(all_len.sum())
        else:
            assert parts is None
    if rank == 0:
        ret["ok"] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_consensus_gloo_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 101, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
