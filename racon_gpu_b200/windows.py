"""Flat (columnar) window batches and the seeded synthetic window generator.

A *window* is what racon calls `racon::Window` (/root/reference/src/window.hpp:71-73): one backbone
segment plus the read layers that overlap it, each layer with an optional PHRED quality string and
a (begin, end) span on the backbone.  The reference keeps a window as vectors of (ptr, len) pairs
and re-packs them per batch (src/cuda/cudabatch.cpp:77-123); here a whole set of windows is ONE
columnar arena so that a batch reaches the GPU as a handful of contiguous copies.

Layout (all numpy, C-contiguous):
    win_seq_off  int64 [W+1]   first sequence index of each window (sequence 0 = backbone)
    seq_off      int64 [S+1]   byte offset of each sequence in `bases` / `weights`
    bases        uint8 [B]     concatenated ASCII bases
    weights      int8  [B]     per-base weight (PHRED - 33), same offsets as `bases`
    has_weights  uint8 [S]     0 => the sequence has no quality string (weight 1 per base)
    begins/ends  int32 [S]     layer span on the backbone (window.cpp:42-63); backbone: (0, 0)
Sequences are stored in ADD order (the order `add_layer` was called in); the processing order is
derived by the host library exactly as window.cpp:78-85 does.

The generator implements SURVEY.md §8(d): uniform random backbone truth, sequence 0 a mutated copy
with weight 0 (the '!' dummy quality racon gives FASTA targets, polisher.cpp:171,392-395), reads
1..D independent mutated copies spanning the whole window, error rate split evenly between
substitutions, insertions and deletions.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class WindowBatch:
    win_seq_off: np.ndarray
    seq_off: np.ndarray
    bases: np.ndarray
    weights: np.ndarray
    has_weights: np.ndarray
    begins: np.ndarray
    ends: np.ndarray

    @property
    def n_windows(self) -> int:
        return int(self.win_seq_off.shape[0] - 1)

    @property
    def n_seqs(self) -> int:
        return int(self.seq_off.shape[0] - 1)

    def window(self, w: int):
        """(sequences, weights-or-None, begins, ends) of window w as Python objects (add order)."""
        s0, s1 = int(self.win_seq_off[w]), int(self.win_seq_off[w + 1])
        seqs, wts = [], []
        for s in range(s0, s1):
            a, b = int(self.seq_off[s]), int(self.seq_off[s + 1])
            seqs.append(self.bases[a:b].tobytes())
            wts.append(self.weights[a:b].copy() if self.has_weights[s] else None)
        return seqs, wts, self.begins[s0:s1].copy(), self.ends[s0:s1].copy()

    def slice(self, w0: int, w1: int) -> "WindowBatch":
        """Windows [w0, w1) as an independent batch (offsets rebased)."""
        s0, s1 = int(self.win_seq_off[w0]), int(self.win_seq_off[w1])
        b0, b1 = int(self.seq_off[s0]), int(self.seq_off[s1])
        return WindowBatch(
            win_seq_off=(self.win_seq_off[w0:w1 + 1] - s0).copy(),
            seq_off=(self.seq_off[s0:s1 + 1] - b0).copy(),
            bases=self.bases[b0:b1].copy(),
            weights=self.weights[b0:b1].copy(),
            has_weights=self.has_weights[s0:s1].copy(),
            begins=self.begins[s0:s1].copy(),
            ends=self.ends[s0:s1].copy(),
        )

    @staticmethod
    def from_lists(windows) -> "WindowBatch":
        """windows: list of lists of (seq: bytes, weights: array|None, begin, end); item 0 = backbone."""
        win_seq_off, seq_off = [0], [0]
        bases, weights, has_w, begins, ends = [], [], [], [], []
        for win in windows:
            for (seq, wt, b, e) in win:
                arr = np.frombuffer(bytes(seq), dtype=np.uint8)
                bases.append(arr)
                if wt is None:
                    weights.append(np.ones(arr.shape[0], dtype=np.int8))
                    has_w.append(0)
                else:
                    wt = np.asarray(wt, dtype=np.int8)
                    assert wt.shape[0] == arr.shape[0]
                    weights.append(wt)
                    has_w.append(1)
                begins.append(b)
                ends.append(e)
                seq_off.append(seq_off[-1] + arr.shape[0])
            win_seq_off.append(len(seq_off) - 1)
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt))
        return WindowBatch(
            win_seq_off=np.asarray(win_seq_off, dtype=np.int64),
            seq_off=np.asarray(seq_off, dtype=np.int64),
            bases=cat(bases, np.uint8),
            weights=cat(weights, np.int8),
            has_weights=np.asarray(has_w, dtype=np.uint8),
            begins=np.asarray(begins, dtype=np.int32),
            ends=np.asarray(ends, dtype=np.int32),
        )


def _mutate_many(truth: np.ndarray, errs: np.ndarray, rng: np.random.Generator):
    """Mutate `truth` (uint8 codes 0..3, shape [W, L]) into len(errs) noisy copies per window.

    Copy c uses total error rate errs[c]: per truth base, with prob e/3 substitute (a different
    base), e/3 delete, e/3 insert one random base after it.  Returns (flat base codes in
    (window, copy, position) order, lengths int64 [W, n_copies]).
    """
    W, L = truth.shape
    C = errs.shape[0]
    # one uint16 decides the edit class, one uint8 supplies the substituted / inserted base
    thr = np.round(errs * 65536.0 / 3.0).astype(np.int64)[None, :, None]
    u = rng.integers(0, 65536, size=(W, C, L), dtype=np.uint16).astype(np.int64)
    r8 = rng.integers(0, 256, size=(W, C, L), dtype=np.uint8)
    t = np.broadcast_to(truth[:, None, :], (W, C, L))
    sub = u < thr
    dele = (u >= thr) & (u < 2 * thr)
    ins = (u >= 2 * thr) & (u < 3 * thr)
    shift = (r8 % 3 + 1).astype(np.uint8)
    emitted = np.where(sub, (t + shift) & 3, t).astype(np.uint8)
    ins_base = ((r8 >> 4) & 3).astype(np.uint8)
    # each truth base emits 0 (deleted), 1, or 2 (base + inserted base) symbols
    count = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int64)
    empty = count.sum(axis=2) == 0  # a sequence must not be empty
    if empty.any():
        count[empty, 0] = 1
    lengths = count.sum(axis=2)
    flat_count = count.reshape(-1)
    start = np.cumsum(flat_count) - flat_count
    out = np.empty(int(flat_count.sum()), dtype=np.uint8)
    keep = flat_count > 0
    out[start[keep]] = emitted.reshape(-1)[keep]
    two = flat_count == 2
    out[start[two] + 1] = ins_base.reshape(-1)[two]
    return out, lengths


def synth_windows(n_windows: int, length: int = 500, depth: int = 32, err: float = 0.15,
                  seed: int = 12345, with_quality: bool = False,
                  backbone_err: float | None = None, chunk: int = 4096) -> WindowBatch:
    """SURVEY.md §8(d) synthetic windows: backbone + `depth` reads => depth+1 sequences per window.

    Config A: length=500, depth=32, err=0.15; B: 1024/64/0.12; C: 500/8/0.05.
    `with_quality`: reads carry PHRED in [5, 25]; otherwise no quality (weight 1).
    The backbone layer is a mutated copy of the truth at `backbone_err` (default: err) and always
    carries weight 0 (racon's '!' dummy quality for FASTA targets).
    Deterministic in (seed, chunk): windows are generated `chunk` at a time from one PCG64 stream.
    """
    rng = np.random.default_rng(seed)
    if backbone_err is None:
        backbone_err = err
    n_per = depth + 1
    errs = np.full(n_per, err, dtype=np.float64)
    errs[0] = backbone_err
    base_parts, len_parts = [], []
    for w0 in range(0, n_windows, chunk):
        nw = min(chunk, n_windows - w0)
        truth = rng.integers(0, 4, size=(nw, length), dtype=np.uint8)
        codes, lens = _mutate_many(truth, errs, rng)
        base_parts.append(_ALPHABET[codes])
        len_parts.append(lens.reshape(-1))
    bases = np.concatenate(base_parts) if base_parts else np.zeros(0, dtype=np.uint8)
    lens = np.concatenate(len_parts) if len_parts else np.zeros(0, dtype=np.int64)
    seq_off = np.zeros(n_windows * n_per + 1, dtype=np.int64)
    np.cumsum(lens, out=seq_off[1:])
    has_w = np.zeros(n_windows * n_per, dtype=np.uint8)
    has_w[0::n_per] = 1
    is_bb_seq = np.zeros(n_windows * n_per, dtype=bool)
    is_bb_seq[0::n_per] = True
    is_bb = np.repeat(is_bb_seq, lens)
    if with_quality:
        weights = rng.integers(5, 26, size=bases.shape[0], dtype=np.int8)
        has_w[:] = 1
    else:
        weights = np.ones(bases.shape[0], dtype=np.int8)
    weights[is_bb] = 0
    begins = np.zeros(n_windows * n_per, dtype=np.int32)
    ends = np.repeat(lens[0::n_per], n_per).astype(np.int32) - 1
    ends[0::n_per] = 0  # backbone position is (0, 0) (window.cpp:38)
    return WindowBatch(
        win_seq_off=np.arange(0, n_windows * n_per + 1, n_per, dtype=np.int64),
        seq_off=seq_off, bases=bases, weights=weights, has_weights=has_w,
        begins=begins, ends=ends)


def edit_distance(a: bytes, b: bytes) -> int:
    """Plain Levenshtein distance (the metric racon's tests use via edlib, test/racon_test.cpp:16-25)."""
    if a == b:
        return 0
    x = np.frombuffer(a, dtype=np.uint8)
    y = np.frombuffer(b, dtype=np.uint8)
    prev = np.arange(y.shape[0] + 1, dtype=np.int32)
    for i in range(x.shape[0]):
        cur = np.empty_like(prev)
        cur[0] = i + 1
        sub = prev[:-1] + (y != x[i])
        dele = prev[1:] + 1
        best = np.minimum(sub, dele)
        # insertion chain: cur[j] = min(best[j-1], cur[j-1] + 1)  -> prefix-min trick
        idx = np.arange(1, y.shape[0] + 1, dtype=np.int32)
        t = np.minimum.accumulate(np.concatenate(([cur[0]], best)) - np.arange(0, y.shape[0] + 1))
        cur[1:] = np.minimum(best, t[1:] + idx)
        prev = cur
    return int(prev[-1])
