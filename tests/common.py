"""Shared helpers for the test-suite: golden fixtures and window builders."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

from racon_gpu_b200.windows import WindowBatch, synth_windows

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M, X, G = 3, -5, -4  # racon scoring named by BASELINE.json


def spoa_sample():
    """(reads, qualities) of spoa's sample.fastq (tests/golden/make_golden.py)."""
    with gzip.open(os.path.join(GOLDEN, "spoa_sample.fastq.gz"), "rt") as fh:
        lines = [l.rstrip("\n") for l in fh]
    reads = [lines[i + 1].encode() for i in range(0, len(lines) - 3, 4)]
    quals = [lines[i + 3].encode() for i in range(0, len(lines) - 3, 4)]
    return reads, quals


def spoa_golden():
    return json.load(open(os.path.join(GOLDEN, "spoa_golden.json")))


def spoa_window(use_quality: bool) -> WindowBatch:
    """spoa's GlobalConsensus test as one window: read 0 is the backbone, the rest full-span layers."""
    reads, quals = spoa_sample()
    L = len(reads[0])
    win = []
    for r, q in zip(reads, quals):
        w = (np.frombuffer(q, dtype=np.uint8).astype(np.int16) - 33).astype(np.int8) if use_quality else None
        win.append((r, w, 0, L - 1))
    win[0] = (win[0][0], win[0][1] if use_quality else None, 0, 0)
    return WindowBatch.from_lists([win])


def ref_fixture(name: str):
    """(WindowBatch, reference consensus untrimmed, reference consensus trimmed) for fixture `name`."""
    z = np.load(os.path.join(GOLDEN, "ref_windows.npz"))
    n, L, D, e1000, q, seed = [int(v) for v in z[name + "_params"]]
    b = synth_windows(n, L, D, e1000 / 1000.0, seed=seed, with_quality=bool(q))
    c0 = z[name + "_cons_0"].tobytes().split(b"\n")
    c1 = z[name + "_cons_1"].tobytes().split(b"\n")
    assert len(c0) == n and len(c1) == n
    return b, c0, c1


def identity_order(batch: WindowBatch) -> np.ndarray:
    order = np.zeros(batch.n_seqs, dtype=np.int32)
    for w in range(batch.n_windows):
        s0, s1 = int(batch.win_seq_off[w]), int(batch.win_seq_off[w + 1])
        order[s0:s1] = np.arange(s1 - s0)
    return order


def partial_span_windows(n: int = 12, length: int = 400, depth: int = 14, seed: int = 77) -> WindowBatch:
    """Windows in which every other layer covers only part of the backbone (window.cpp:96-103:
    such layers are aligned to a subgraph), like real racon windows at read ends."""
    rng = np.random.default_rng(5)
    b = synth_windows(n, length, depth, 0.1, seed=seed)
    wins = []
    for w in range(b.n_windows):
        seqs, wts, _, _ = b.window(w)
        L = len(seqs[0])
        win = [(seqs[0], wts[0], 0, 0)]
        for i in range(1, len(seqs)):
            if i % 2 == 0:
                lo, hi = sorted(rng.integers(0, L, size=2).tolist())
                if hi - lo < 40:
                    lo, hi = 10, L - 10
                s = seqs[i][int(lo / L * len(seqs[i])):int(hi / L * len(seqs[i]))]
                win.append((s, None, lo, hi))
            else:
                win.append((seqs[i], None, 0, L - 1))
        wins.append(win)
    return WindowBatch.from_lists(wins)


def awkward_windows(m: int, x: int, g: int, n: int = 36) -> WindowBatch:
    """Many small, awkward windows: short and long layers, partial spans, with and without qualities, deep and
    shallow, up to 30 % error (used by the emulation and the GPU parity tests with several scoring schemes)."""
    rng = np.random.default_rng(1000 + 7 * m - x - 13 * g)
    wins = []
    for _ in range(n):
        L = int(rng.integers(30, 260))
        D = int(rng.integers(2, 24))
        err = float(rng.uniform(0.0, 0.3))
        b = synth_windows(1, L, D, err, seed=int(rng.integers(1 << 30)), with_quality=bool(rng.integers(2)))
        seqs, wts, _, _ = b.window(0)
        Lb = len(seqs[0])
        win = [(seqs[0], wts[0], 0, 0)]
        for i in range(1, len(seqs)):
            if rng.random() < 0.35 and Lb > 60:  # a layer covering part of the backbone only
                lo, hi = sorted(rng.integers(0, Lb, size=2).tolist())
                if hi - lo < 20:
                    lo, hi = 5, Lb - 5
                a, z = int(lo / Lb * len(seqs[i])), int(hi / Lb * len(seqs[i]))
                if z - a >= 2:
                    win.append((seqs[i][a:z], None if wts[i] is None else wts[i][a:z], lo, hi))
                    continue
            win.append((seqs[i], wts[i], 0, Lb - 1))
        wins.append(win)
    return WindowBatch.from_lists(wins)
