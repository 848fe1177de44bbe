"""Multi-GPU plumbing: windows are independent, so ranks own contiguous window ranges and the only
collective is the final gather of (length, consensus) to rank 0.

One process per GPU (`torch.distributed`, backend nccl on GPUs / gloo on CPU for the tests).  The
reference is one process with one host thread per (device, batch), src/cuda/cudapolisher.cpp:228-240,
336-345; that form is b200poa_polisher_create(devices=[...]).

`ConsensusGather` is the collective of the path.  What travels is the COMPACT consensus (sum of the lengths,
~0.5 KB per window, not rows padded to the engine's 2 KB maximum), staged through pinned host memory, issued on a
side CUDA stream and awaited one step later, so that it overlaps the next step's polishing; nothing is
concatenated on the critical path (rank 0 keeps one pinned block per rank).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api


def shard_range(n_windows: int, rank: int, world: int):
    """Contiguous, balanced split: rank r owns [lo, hi)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ConsensusGather:
    """Gathers every rank's (lengths, compact consensus bytes) to rank 0, in rank order.

        g = ConsensusGather(device, n_local_windows)
        g.start(cons, clen)      # cons [w_r, stride] uint8 rows valid up to clen[w]; returns at once
        ...                      # the next step's work
        parts = g.wait()         # rank 0: list over ranks of (lens int32 [w_r], flat uint8 [sum lens]); else None

    `assemble(parts)` turns that into one (flat, offsets) pair when a caller wants it (off the critical path).
    """

    def __init__(self, device: torch.device, n_local: int):
        self.device = device
        self.cuda = device.type == "cuda"
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_local = int(n_local)
        counts = torch.tensor([self.n_local], dtype=torch.int64, device=device)
        allc = torch.zeros(self.world, dtype=torch.int64, device=device)
        if self.world > 1:
            dist.all_gather_into_tensor(allc, counts)
        else:
            allc[0] = self.n_local
        self.counts = [int(c) for c in allc.cpu().tolist()]  # one-time setup, not on the per-step path
        self.n_max = max(self.counts + [1])
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = None
        self.cap = 0
        self._pending = None
        pin = self.cuda
        self.h_lens = torch.zeros(self.n_max, dtype=torch.int32, pin_memory=pin)
        self.d_lens = torch.zeros(self.n_max, dtype=torch.int32, device=device)
        self.h_flat = None
        self.d_flat = None
        self.d_sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.g_lens = ([torch.zeros(self.n_max, dtype=torch.int32, device=device) for _ in range(self.world)]
                       if self.rank == 0 else None)
        self.h_all_lens = torch.zeros((self.world, self.n_max), dtype=torch.int32, pin_memory=pin) if self.rank == 0 else None
        self.g_flat = None
        self.h_all_flat = None

    def _ensure(self, attr, n, **kw):
        t = getattr(self, attr)
        if t is None or t.shape[-1] < n:
            n = int(n * 1.25) + 4096
            setattr(self, attr, torch.zeros(n, dtype=torch.uint8, **kw))
        return getattr(self, attr)

    def start(self, cons: np.ndarray, clen: np.ndarray):
        assert self._pending is None, "wait() for the previous gather first"
        W = cons.shape[0]
        assert W == self.n_local
        # pack this rank's rows back to back, straight into pinned memory
        total = int(np.minimum(clen, cons.shape[1]).sum())
        h_flat = self._ensure("h_flat", total, pin_memory=self.cuda)
        flat_np, _ = api.compact_rows(cons, clen, h_flat.numpy())
        self.h_lens[:W] = torch.from_numpy(np.ascontiguousarray(clen, dtype=np.int32))
        if self.world == 1:
            self._pending = ("local", total)
            return
        ctx = torch.cuda.stream(self.stream) if self.cuda else _Null()
        with ctx:
            self.d_lens.copy_(self.h_lens, non_blocking=True)
            my = torch.tensor([total], dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(self.d_sizes, my)
            sizes = [int(x) for x in self.d_sizes.cpu().tolist()]  # 8 numbers; the only host wait of start()
            cap = max(sizes + [1])
            d_flat = self._ensure("d_flat", cap, device=self.device)
            d_flat[:total].copy_(h_flat[:total], non_blocking=True)
            if self.rank == 0:
                if self.g_flat is None or self.g_flat[0].shape[0] < cap:
                    n = int(cap * 1.25) + 4096
                    self.g_flat = [torch.zeros(n, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
                    self.h_all_flat = torch.zeros((self.world, n), dtype=torch.uint8, pin_memory=self.cuda)
                dist.gather(self.d_lens, self.g_lens, dst=0)
                dist.gather(d_flat[:cap], [g[:cap] for g in self.g_flat], dst=0)
                for r in range(self.world):  # device -> pinned host, asynchronous on the side stream
                    self.h_all_lens[r].copy_(self.g_lens[r], non_blocking=True)
                    self.h_all_flat[r, :sizes[r]].copy_(self.g_flat[r][:sizes[r]], non_blocking=True)
            else:
                dist.gather(self.d_lens, None, dst=0)
                dist.gather(d_flat[:cap], None, dst=0)
            if self.cuda:
                self.done = torch.cuda.Event()
                self.done.record(self.stream)
        self._pending = ("dist", sizes)

    def wait(self):
        if self._pending is None:
            return None
        kind, info = self._pending
        self._pending = None
        if kind == "local":
            return [(self.h_lens[:self.n_local].numpy(), self.h_flat[:info].numpy())]
        if self.cuda:
            self.done.synchronize()
        if self.rank != 0:
            return None
        return [(self.h_all_lens[r, :self.counts[r]].numpy(), self.h_all_flat[r, :info[r]].numpy())
                for r in range(self.world)]


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def assemble(parts):
    """(flat uint8, offsets int64 [W+1]) of the whole job from ConsensusGather.wait()'s per-rank parts."""
    lens = np.concatenate([p[0] for p in parts]).astype(np.int64)
    off = np.zeros(lens.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return np.concatenate([p[1] for p in parts]), off


def gather_consensus(cons: np.ndarray, clen: np.ndarray, device: torch.device):
    """One-shot form: returns (flat, offsets) of the whole job on rank 0 and (None, None) elsewhere."""
    g = ConsensusGather(device, cons.shape[0])
    g.start(cons, clen)
    parts = g.wait()
    return assemble(parts) if parts is not None else (None, None)
