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
    """Gather every rank's padded consensus rows [w_r, stride] and lengths to rank 0 in rank order.

    Rows are fixed-stride so this is one gather of a [W_max, stride] uint8 tensor plus one of the
    lengths.  Returns (cons, clen) of the whole job on rank 0 and (None, None) elsewhere.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return cons, clen
    rank = dist.get_rank()
    n_local = torch.tensor([cons.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    w_max, stride = max(counts), cons.shape[1]
    pad_c = torch.zeros((w_max, stride), dtype=torch.uint8, device=device)
    pad_l = torch.zeros((w_max,), dtype=torch.int32, device=device)
    pad_c[:cons.shape[0]] = torch.from_numpy(cons).to(device, non_blocking=True)
    pad_l[:clen.shape[0]] = torch.from_numpy(clen).to(device, non_blocking=True)
    out_c = [torch.empty_like(pad_c) for _ in range(world)] if rank == 0 else None
    out_l = [torch.empty_like(pad_l) for _ in range(world)] if rank == 0 else None
    dist.gather(pad_c, out_c, dst=0)
    dist.gather(pad_l, out_l, dst=0)
    if rank != 0:
        return None, None
    all_c = torch.cat([out_c[r][:counts[r]] for r in range(world)]).cpu().numpy()
    all_l = torch.cat([out_l[r][:counts[r]] for r in range(world)]).cpu().numpy()
    return all_c, all_l
