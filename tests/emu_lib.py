"""ctypes loader of the host-flavour engine build (tests/emu/libpoa_emu.so; TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import _flat_args, _p

HERE = os.path.dirname(os.path.abspath(__file__))


class Emu:
    def __init__(self):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu")], check=True,
                       stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(os.path.join(HERE, "emu", "libpoa_emu.so"))

    def polish(self, b, order, m, x, g, max_nodes=4096, max_edges=24576, max_len=1023, band=0,
               serial_topsort=False, threads=8, stride=4096, msa_cap=0):
        """Untrimmed consensus, coverage, status and DP cell count of the emulated engine."""
        W = b.n_windows
        cons = np.zeros((W, stride), dtype=np.uint8)
        cov = np.zeros((W, stride), dtype=np.uint16)
        clen = np.zeros(W, dtype=np.int32)
        st = np.zeros(W, dtype=np.int32)
        cells = C.c_int64(0)
        self.last_trim = np.zeros(W, dtype=np.int32)  # device-side trim interval: first | last << 16
        order = np.ascontiguousarray(order, dtype=np.int32)
        a = _flat_args(b)
        self.lib.emu_polish_windows(
            a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], _p(order, C.c_int32), C.c_int32(m), C.c_int32(x), C.c_int32(g),
            C.c_int32(max_nodes), C.c_int32(max_edges), C.c_int32(max_len), C.c_int32(band),
            C.c_int32(int(serial_topsort)), C.c_int32(threads), _p(cons, C.c_uint8), _p(cov, C.c_uint16),
            C.c_int32(stride), _p(clen, C.c_int32), _p(st, C.c_int32), None, None, C.byref(cells),
            _p(self.last_trim, C.c_int32), *self._msa_args(W, msa_cap))
        self.last_msa = None
        if msa_cap:  # per window: list of rows (processing order), or the status when the MSA failed
            arena, off, cols, mst = self._msa
            nseq = np.diff(b.win_seq_off)
            self.last_msa = [[arena[off[w] + k * cols[w]:off[w] + (k + 1) * cols[w]].tobytes() for k in range(int(nseq[w]))]
                             if mst[w] == 0 else int(mst[w]) for w in range(W)]
        return ([cons[w, :clen[w]].tobytes() for w in range(W)], [cov[w, :clen[w]].copy() for w in range(W)],
                st, cells.value)

    def _msa_args(self, W, msa_cap):
        if not msa_cap:
            self._msa = None
            return (None, C.c_int64(0), None, None, None)
        self._msa = (np.zeros(msa_cap, dtype=np.uint8), np.zeros(W, dtype=np.int64), np.zeros(W, dtype=np.int32),
                     np.zeros(W, dtype=np.int32))
        a, o, c, s = self._msa
        return (_p(a, C.c_uint8), C.c_int64(msa_cap), _p(o, C.c_int64), _p(c, C.c_int32), _p(s, C.c_int32))


class EmuAligner:
    """Host-flavour build of the overlap aligner (tests/emu/libaln_emu.so)."""

    def __init__(self):
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu")], check=True, stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(os.path.join(HERE, "emu", "libaln_emu.so"))
        self.lib.emu_align.restype = C.c_int64

    def align(self, q: bytes, t: bytes, q_first: int = 0, t_begin: int = 0, window_length: int = 0, guess: int = -1):
        """(ops, edit distance, recursion depth, leaves); self.cigar = the CIGAR formed from the engine's runs"""
        ops = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
        score, levels, leaves = C.c_int32(-1), C.c_int32(0), C.c_int32(0)
        cap = 4 * (len(q) + len(t)) + 16
        cigar = C.create_string_buffer(cap)
        bp = np.zeros(4 * (len(t) // max(window_length, 1) + 3), dtype=np.uint32)
        bp_count = C.c_int32(0)
        n = self.lib.emu_align(C.c_char_p(q), C.c_int32(len(q)), C.c_char_p(t), C.c_int32(len(t)), _p(ops, C.c_uint8),
                               C.byref(score), C.byref(levels), C.byref(leaves), cigar, C.c_int64(cap),
                               C.c_int32(q_first), C.c_int32(t_begin), C.c_int32(window_length), _p(bp, C.c_uint32),
                               C.byref(bp_count), C.c_int32(guess))
        self.breaking_points = bp[:2 * bp_count.value].reshape(-1, 2).copy()  # (t, q) pairs, window_length > 0 only
        assert n >= 0, {-1: "inconsistent split", -2: "list overflow", -3: "runs do not spell the operations"}.get(n, n)
        self.cigar = cigar.value
        return ops[:n].copy(), int(score.value), int(levels.value), int(leaves.value)
