"""CPU: the engine's graph phases (racon_gpu_b200/csrc/poa_core.cuh, host flavour) against the oracle.

The same source compiles into the CUDA kernel; here POA_LANES loops over 32 lanes so the parallel
add_alignment, the per-root topological sort, the ballot traceback, the band definition and the
consensus are checked bit-for-bit without a GPU.
"""
import numpy as np
import pytest

from common import G, M, X, identity_order, spoa_golden, spoa_window
from emu_lib import Emu
from racon_gpu_b200 import api
from racon_gpu_b200.windows import WindowBatch, edit_distance, synth_windows


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.mark.parametrize("cfg", [(24, 500, 8, 0.05, False), (24, 500, 32, 0.15, False), (16, 400, 20, 0.12, True),
                                 (4, 900, 40, 0.12, False), (12, 300, 60, 0.25, False)])
@pytest.mark.parametrize("serial", [True, False])
def test_full_band_is_bit_exact(emu, oracle, cfg, serial):
    n, L, D, e, q = cfg
    b = synth_windows(n, L, D, e, seed=31, with_quality=q)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=8)
    ec, ecov, st, _ = emu.polish(b, order, M, X, G, band=0, serial_topsort=serial)
    assert (st == 0).all()
    assert ec == oc
    assert all((a == c).all() for a, c in zip(ecov, ocov))


def test_per_root_topsort_equals_serial_dfs_on_deep_windows(emu):
    """Deep, noisy, quality-weighted windows: the decomposed sort must equal spoa's DFS order."""
    b = synth_windows(8, 250, 150, 0.2, seed=8, with_quality=True)
    order = api.processing_order(b)
    a = emu.polish(b, order, M, X, G, serial_topsort=True)
    c = emu.polish(b, order, M, X, G, serial_topsort=False)
    assert (a[2] == 0).all() and a[0] == c[0]
    assert all((x == y).all() for x, y in zip(a[1], c[1]))


def test_spoa_known_answers(emu):
    gold = spoa_golden()
    sc = gold["scoring"]
    for use_q, key in ((False, "GlobalConsensus"), (True, "GlobalConsensusWithQualities")):
        b = spoa_window(use_q)
        ec, _, st, _ = emu.polish(b, identity_order(b), sc["m"], sc["x"], sc["g"])
        assert st[0] == 0 and ec[0].decode() == gold[key]


def test_partial_span_layers_go_through_the_subgraph_path(emu, oracle):
    """window.cpp:96-103 / graph.cpp:592-683: layers that do not span the window."""
    from common import partial_span_windows
    pb = partial_span_windows()
    order = api.processing_order(pb)
    oc, ocov, _ = oracle.polish(pb, order, M, X, G, tgs=False, trim=False, threads=8)
    for band in (0, 256):
        for serial in (True, False):
            ec, ecov, st, _ = emu.polish(pb, order, M, X, G, band=band, serial_topsort=serial)
            assert (st == 0).all() and ec == oc
            assert all((a == c).all() for a, c in zip(ecov, ocov))


def test_static_band_256_tolerance(emu, oracle):
    """Banded mode is compared with the UNBANDED oracle.  Stated tolerance: >= 99% of
    windows identical, per-window edit distance <= 2."""
    for (n, L, D, e) in [(48, 500, 32, 0.15), (6, 900, 64, 0.12)]:
        b = synth_windows(n, L, D, e, seed=17)
        order = api.processing_order(b)
        oc, _, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=8)
        ec, _, st, cells = emu.polish(b, order, M, X, G, band=256)
        assert (st == 0).all()
        d = [edit_distance(a, c) for a, c in zip(oc, ec)]
        assert sum(x == 0 for x in d) >= 0.99 * n and max(d) <= 2


def test_limits_report_status_instead_of_crashing(emu):
    b = synth_windows(4, 500, 32, 0.15, seed=3)
    order = api.processing_order(b)
    _, _, st, _ = emu.polish(b, order, M, X, G, max_nodes=700, max_edges=8000)
    assert (st == 4).all()  # node_count_exceeded_maximum_graph_size
    # (edge_count_exceeded is no longer reachable on its own: the edge pool is 6 x the node capacity by construction,
    #  poa_edge_capacity(), and a graph runs out of nodes long before it averages six in-edges per node)
    _, _, st, _ = emu.polish(b, order, M, X, G, max_len=400)
    assert (st == 2).all()  # exceeded_maximum_sequence_size


def test_long_window_stress_with_a_larger_sequence_limit(emu, oracle):
    """BASELINE config 5 (1024 bp x 64 reads, 12 %): half the reads exceed racon's 1023-base BatchConfig; with the
    limit raised (graph limits scale with it, batch.cu:34-71) full band is bit-exact and band 256 matches."""
    b = synth_windows(4, 1024, 64, 0.12, seed=5)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=8)
    ec, ecov, st, _ = emu.polish(b, order, M, X, G, band=0, max_len=1279, max_nodes=3840, max_edges=6 * 3840)
    assert (st == 0).all() and ec == oc and all((a == c).all() for a, c in zip(ecov, ocov))
    bc, _, st, _ = emu.polish(b, order, M, X, G, band=256, max_len=1279, max_nodes=5116, max_edges=6 * 5116)
    assert (st == 0).all() and sum(a == c for a, c in zip(bc, oc)) >= 3


@pytest.mark.parametrize("scoring", [(3, -5, -4), (5, -4, -8), (1, -1, -1), (2, -3, -2)])
def test_randomized_small_windows_match_the_oracle(emu, oracle, scoring):
    """Many small, awkward windows (short and long layers, partial spans, with and without qualities, deep and
    shallow, high error) under several scoring schemes: consensus AND coverage equal the oracle, both sorts."""
    from common import awkward_windows
    m, x, g = scoring
    b = awkward_windows(m, x, g)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, m, x, g, tgs=False, trim=False, threads=8)
    for serial in (False, True):
        ec, ecov, st, _ = emu.polish(b, order, m, x, g, band=0, serial_topsort=serial)
        assert (st == 0).all() and ec == oc
        assert all((a == c).all() for a, c in zip(ecov, ocov))


def deletion_windows(n=12, seed=5):
    """Windows in which every other read lacks 150-260 consecutive bases: their alignments leave a 256-column band."""
    import numpy as np
    from racon_gpu_b200.windows import WindowBatch, synth_windows
    rng = np.random.default_rng(3)
    b = synth_windows(n, 900, 16, 0.08, seed=seed)
    wins = []
    for w in range(b.n_windows):
        seqs, wts, _, _ = b.window(w)
        L = len(seqs[0])
        win = [(seqs[0], wts[0], 0, 0)]
        for i in range(1, len(seqs)):
            s = seqs[i]
            if i % 2 == 0:
                a, d = int(rng.integers(100, 300)), int(rng.integers(150, 260))
                s = s[:a] + s[a + d:]
            win.append((s, None, 0, L - 1))
        wins.append(win)
    return WindowBatch.from_lists(wins)


def test_adaptive_band_recovers_what_the_static_band_loses(oracle):
    """adaptive band = the static band plus cudapoa's retry protocol (a traceback near the band edge
    re-aligns the read with twice the width).  On windows with long deletions the static band drifts from the unbanded
    oracle, the adaptive band is exact at a fraction of the full band's cells."""
    from common import identity_order
    from emu_lib import Emu
    from racon_gpu_b200.windows import edit_distance
    b = deletion_windows()
    order = identity_order(b)
    oc, ocov, _ = oracle.polish(b, order, 3, -5, -4, tgs=False, trim=False, threads=8, stride=8192)
    emu = Emu()
    sc, _, sst, scells = emu.polish(b, order, 3, -5, -4, max_nodes=4092, max_edges=24000, band=256, stride=8192)
    ac, acov, ast, acells = emu.polish(b, order, 3, -5, -4, max_nodes=4092, max_edges=24000, band=-256, stride=8192)
    _, _, _, fcells = emu.polish(b, order, 3, -5, -4, max_nodes=4092, max_edges=24000, band=0, stride=8192)
    assert (ast == 0).all() and ac == oc and all((a == c).all() for a, c in zip(acov, ocov))
    assert sum(edit_distance(a, c) for a, c in zip(sc, oc)) > 0       # the case is one the static band gets wrong
    assert scells < acells < 0.5 * fcells


def test_int32_cells_where_int16_cannot_hold_the_alignment(emu, oracle):
    """reads whose alignment does not provably fit int16 (score_range_ok) are filled and traced with 32-bit
    cells, the switch spoa makes (simd_alignment_engine.cpp:668-673).  Forced on ordinary windows the 32-bit path must
    be exact; an extreme scoring scheme takes it by itself."""
    b = synth_windows(12, 300, 16, 0.15, seed=9, with_quality=True)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=8)
    ec, ecov, st, _ = emu.polish(b, order, M, X, G, band=0, serial_topsort=2)  # bit 1: force 32-bit cells
    assert (st == 0).all() and ec == oc and all((a == c).all() for a, c in zip(ecov, ocov))
    bc, _, st, _ = emu.polish(b, order, M, X, G, band=256, serial_topsort=2)
    assert (st == 0).all() and bc == oc
    pb = __import__("common").partial_span_windows()
    po = api.processing_order(pb)
    oc, ocov, _ = oracle.polish(pb, po, M, X, G, tgs=False, trim=False, threads=8)
    ec, ecov, st, _ = emu.polish(pb, po, M, X, G, band=0, serial_topsort=2)
    assert (st == 0).all() and ec == oc and all((a == c).all() for a, c in zip(ecov, ocov))
    b = synth_windows(4, 500, 32, 0.15, seed=3)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, 120, -120, -120, tgs=False, trim=False, threads=8)
    ec, _, st, _ = emu.polish(b, order, 120, -120, -120)
    assert (st == 0).all() and ec == oc
