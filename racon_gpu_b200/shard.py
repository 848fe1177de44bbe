"""Multi-GPU plumbing: windows are independent, so ranks own contiguous window ranges and the only
collective is the final gather of (length, consensus) rows to rank 0.

One process per GPU (`torch.distributed`, backend nccl on GPUs / gloo on CPU for the tests).  The
reference is one process with one host thread per (device, batch), src/cuda/cudapolisher.cpp:228-240,
336-345; that form is b200poa_polish_windows(devices=[...]).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_windows: int, rank: int, world: int):
    """Contiguous, balanced split: rank r owns [lo, hi)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_consensus(cons: np.ndarray, clen: np.ndarray, device: torch.device):
    """Gather every rank's consensus rows and lengths to rank 0 in rank order (the one collective of the
    path: windows are independent, cudapolisher.cpp:228-345 has no exchange step).

    `cons` is [w_r, stride] with row w valid up to clen[w].  Only the first max(clen) columns travel (the
    rows are padded to the engine's maximum consensus length, ~10x the real one).  Returns (cons, clen) of
    the whole job on rank 0 -- rows [W, max_len], bytes past a row's length unspecified -- and (None, None)
    elsewhere.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return cons, clen
    rank = dist.get_rank()
    local_w = int(clen.max()) if clen.size else 0
    meta = torch.tensor([cons.shape[0], local_w], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    counts = [int(m[0].item()) for m in metas]
    w_max, width = max(counts), max(1, max(int(m[1].item()) for m in metas))
    pad_c = torch.zeros((w_max, width), dtype=torch.uint8, device=device)
    pad_l = torch.zeros((w_max,), dtype=torch.int32, device=device)
    take = min(width, cons.shape[1])
    pad_c[:cons.shape[0], :take] = torch.from_numpy(np.ascontiguousarray(cons[:, :take])).to(device, non_blocking=True)
    pad_l[:clen.shape[0]] = torch.from_numpy(np.ascontiguousarray(clen, dtype=np.int32)).to(device, non_blocking=True)
    out_c = [torch.empty_like(pad_c) for _ in range(world)] if rank == 0 else None
    out_l = [torch.empty_like(pad_l) for _ in range(world)] if rank == 0 else None
    dist.gather(pad_c, out_c, dst=0)
    dist.gather(pad_l, out_l, dst=0)
    if rank != 0:
        return None, None
    all_c = torch.cat([out_c[r][:counts[r]] for r in range(world)]).cpu().numpy()
    all_l = torch.cat([out_l[r][:counts[r]] for r in range(world)]).cpu().numpy()
    return all_c, all_l
