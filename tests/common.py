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


# ---------------------------------------------------------------------------------------------------
# Real racon windows (tests/golden/make_lambda_golden.py): the lambda-phage sample of the reference's own
# end-to-end tests and the 67 deep cudapoa sample windows.
# ---------------------------------------------------------------------------------------------------
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def _unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    c = np.stack([packed & 3, (packed >> 2) & 3, (packed >> 4) & 3, (packed >> 6) & 3], axis=1).reshape(-1)[:n]
    return _ACGT[c]


def _batch_from_columns(bases, seq_len, win_nseq, qual=None, has_q=None, begin=None, end=None) -> WindowBatch:
    S = seq_len.shape[0]
    seq_off = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(seq_len, out=seq_off[1:])
    win_seq_off = np.zeros(win_nseq.shape[0] + 1, dtype=np.int64)
    np.cumsum(win_nseq, out=win_seq_off[1:])
    weights = np.ones(bases.shape[0], dtype=np.int8)
    if has_q is None:
        has_q = np.zeros(S, dtype=np.uint8)
    if qual is not None and qual.shape[0]:
        mask = np.repeat(has_q.astype(bool), seq_len)
        weights[mask] = (qual.astype(np.int16) - 33).astype(np.int8)
    if begin is None:  # all layers span the window (cudapoa sample windows carry no positions)
        bb_len = np.repeat(seq_len[win_seq_off[:-1]], win_nseq)
        begin = np.zeros(S, dtype=np.int32)
        end = (bb_len - 1).astype(np.int32)
        end[win_seq_off[:-1]] = 0
    return WindowBatch(win_seq_off=win_seq_off, seq_off=seq_off, bases=np.ascontiguousarray(bases), weights=weights,
                       has_weights=np.ascontiguousarray(has_q, dtype=np.uint8),
                       begins=np.ascontiguousarray(begin, dtype=np.int32), ends=np.ascontiguousarray(end, dtype=np.int32))


def _split(flat: np.ndarray, lens: np.ndarray):
    off = np.concatenate([[0], np.cumsum(lens)])
    return [flat[off[i]:off[i + 1]].tobytes() for i in range(lens.shape[0])]


def lambda_fixture(case: str):
    """Real racon windows of the lambda-phage sample.  case in {fastq_500, fasta_500, fastq_1000}.
    Returns (WindowBatch in ADD order, racon CPU consensus per window, racon CPU status per window, params dict)."""
    z = np.load(os.path.join(GOLDEN, "lambda_windows.npz"))
    g = lambda k: z[f"{case}_{k}"]
    seq_len = g("seq_len")
    bases = _unpack2(g("bases"), int(seq_len.sum()))
    b = _batch_from_columns(bases, seq_len, g("win_nseq"), g("qual"), g("has_q"), g("begin"), g("end"))
    wl, tgs, trim, m, x, gap, ed, clen = [int(v) for v in g("params")]
    params = {"window_length": wl, "tgs": bool(tgs), "trim": bool(trim), "m": m, "x": x, "g": gap,
              "edit_distance": ed, "contig_length": clen}
    return b, _split(g("cons"), g("cons_len")), g("polished").astype(bool), params


def lambda_reference() -> bytes:
    z = np.load(os.path.join(GOLDEN, "lambda_windows.npz"))
    return _unpack2(z["reference"], int(z["reference_len"][0])).tobytes()


def reverse_complement(s: bytes) -> bytes:
    return _COMP[np.frombuffer(s, dtype=np.uint8)][::-1].tobytes()


def contig_edit_distance(oracle, window_consensus, reference: bytes) -> int:
    """test/racon_test.cpp:98-107: edit distance of the reverse complement of the stitched contig to the reference."""
    import ctypes as C
    q = reverse_complement(b"".join(window_consensus))
    fn = oracle.lib.poa_oracle_edit_distance
    fn.restype = C.c_int64
    return int(fn(C.c_char_p(q), C.c_int64(len(q)), C.c_char_p(reference), C.c_int64(len(reference))))


def cudapoa_fixture():
    """The 67 deep windows of cudapoa's sample-windows.txt (file order = processing order, weight 1, full span).
    Returns (WindowBatch, reference consensus list, reference coverage list)."""
    z = np.load(os.path.join(GOLDEN, "cudapoa_windows.npz"))
    seq_len = z["seq_len"]
    bases = _unpack2(z["bases"], int(seq_len.sum()))
    b = _batch_from_columns(bases, seq_len, z["win_nseq"])
    cons = _split(z["cons"], z["cons_len"])
    off = np.concatenate([[0], np.cumsum(z["cons_len"])])
    cov = [z["cov"][off[i]:off[i + 1]] for i in range(len(cons))]
    return b, cons, cov


def overlap_fixture():
    """The 181 real overlap alignment inputs of the lambda-phage sample (tests/golden/make_overlap_golden.py).
    Returns a list of dicts {q, t, score, n_ops, cigar_sha} (edlib's result as racon's CPU path obtains it)."""
    z = np.load(os.path.join(GOLDEN, "lambda_overlaps.npz"))
    ql, tl = z["q_len"], z["t_len"]
    qs = _split(_unpack2(z["q_bases"], int(ql.sum())), ql)
    ts = _split(_unpack2(z["t_bases"], int(tl.sum())), tl)
    meta = [l.split() for l in z["cigar_sha"].tobytes().decode().strip().split("\n")]
    bo = np.concatenate([[0], np.cumsum(z["bp_count"])]) * 2
    return [{"q": qs[i], "t": ts[i], "score": int(z["score"][i]), "n_ops": int(meta[i][0]), "cigar_sha": meta[i][1],
             "q_first": int(z["q_first"][i]), "t_begin": int(z["t_begin"][i]),
             "bp": z["bp"][bo[i]:bo[i + 1]].reshape(-1, 2)}  # racon's breaking points at window length 500
            for i in range(len(qs))]


def random_pairs(seed: int, shapes):
    """(q, t) pairs: t random over ACGT, q a mutated copy (substitutions, insertions, deletions at rate e)."""
    rng = np.random.default_rng(seed)
    out = []
    for n, e in shapes:
        t = rng.choice(_ACGT, size=n)
        keep = rng.random(n)
        parts = []
        for k in range(n):
            x = keep[k]
            if x < e / 3:
                parts.append(_ACGT[rng.integers(4)])
            elif x < 2 * e / 3:
                continue
            elif x < e:
                parts.append(t[k])
                parts.append(_ACGT[rng.integers(4)])
            else:
                parts.append(t[k])
        q = np.asarray(parts, dtype=np.uint8).tobytes() if parts else b"A"
        out.append((q, t.tobytes()))
    return out
