"""GPU (-m gpu): the CUDA engine, called through the C ABI, against the oracle and the golden fixtures."""
import ctypes as C

import numpy as np
import pytest

from common import G, M, X, identity_order, ref_fixture, spoa_golden, spoa_window
from racon_gpu_b200 import api
from racon_gpu_b200.windows import WindowBatch, edit_distance, synth_windows

pytestmark = pytest.mark.gpu
MEM = 6 << 30


def gpu_untrimmed(b, banded=False, mem=MEM, **kw):
    pb = api.PoaBatch(max_gpu_mem=mem, banded=banded, **kw)
    n, _ = pb.add_windows(b)
    assert n == b.n_windows
    pb.generate_poa()
    out = pb.get_consensus()
    info = pb.info()
    pb.close()
    assert info["kernel_launches"] == 1
    return out


@pytest.mark.parametrize("cfg", [(64, 500, 8, 0.05, False), (96, 500, 32, 0.15, False), (48, 400, 20, 0.12, True),
                                 (8, 900, 64, 0.12, False), (16, 300, 120, 0.2, True)])
def test_full_band_bit_exact_vs_oracle(oracle, cfg):
    n, L, D, e, q = cfg
    b = synth_windows(n, L, D, e, seed=41, with_quality=q)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16)
    gc, gcov, st = gpu_untrimmed(b)
    assert (st == 0).all()
    assert gc == oc
    assert all((a == c).all() for a, c in zip(gcov, ocov))


@pytest.mark.parametrize("name", ["A", "C", "Q", "B"])
@pytest.mark.parametrize("via_adapter", [False, True])
def test_reference_fixtures_through_the_polisher_api(name, via_adapter):
    """Whole job with host buffers against the committed outputs of the unmodified reference."""
    b, ref_untrimmed, ref_trimmed = ref_fixture(name)
    cons, clen, pol, status, _ = api.polish_windows(b, M, X, G, tgs=True, trim=True, mem_per_batch=MEM,
                                                    via_adapter=via_adapter)
    assert pol.all()
    assert api.consensus_list(cons, clen) == ref_trimmed
    cons, clen, pol, status, _ = api.polish_windows(b, M, X, G, tgs=False, trim=False, mem_per_batch=MEM,
                                                    via_adapter=via_adapter)
    assert api.consensus_list(cons, clen) == ref_untrimmed


def test_spoa_known_answers_through_add_group():
    """vendor/spoa/test/spoa_test.cpp:220-238,283-301 (5/-4/-8), entries in file order."""
    gold = spoa_golden()
    sc = gold["scoring"]
    for use_q, key in ((False, "GlobalConsensus"), (True, "GlobalConsensusWithQualities")):
        b = spoa_window(use_q)
        seqs, wts, _, _ = b.window(0)
        pb = api.PoaBatch(max_gpu_mem=MEM, gap=sc["g"], mismatch=sc["x"], match=sc["m"])
        st, per = pb.add_poa_group([(s, w) for s, w in zip(seqs, wts)])
        assert st == 0 and all(p == 0 for p in per)
        pb.generate_poa()
        cons, _, status = pb.get_consensus()
        pb.close()
        assert status[0] == 0 and cons[0].decode() == gold[key]


def test_static_band_tolerance_vs_unbanded_oracle(oracle):
    """BASELINE config 2 shape (500 bp x 32, 15%, band 256): >= 99% identical, edit distance <= 2."""
    b = synth_windows(256, 500, 32, 0.15, seed=43)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16)
    gc, _, st = gpu_untrimmed(b, banded=True)
    assert (st == 0).all()
    d = [edit_distance(a, c) for a, c in zip(oc, gc)]
    assert sum(x == 0 for x in d) >= 0.99 * len(d) and max(d) <= 2
    # long-window stress shape (1024 x 64, 12%): the band is what makes it fit; same tolerance
    b = synth_windows(8, 900, 64, 0.12, seed=44)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16)
    gc, _, st = gpu_untrimmed(b, banded=True)
    d = [edit_distance(a, c) for a, c in zip(oc, gc)]
    assert (st == 0).all() and max(d) <= 2


def test_banded_gpu_equals_banded_emulation():
    """The CUDA fill and its scalar twin must produce the same banded result bit for bit."""
    from emu_lib import Emu
    b = synth_windows(32, 700, 24, 0.15, seed=45)
    order = api.processing_order(b)
    ec, ecov, est, _ = Emu().polish(b, order, M, X, G, band=256, max_nodes=4092)
    gc, gcov, st = gpu_untrimmed(b, banded=True)
    assert (st == est).all() and gc == ec
    assert all((a == c).all() for a, c in zip(gcov, ecov))


def test_edge_cases_follow_the_reference_contract():
    lib = api.load_library()
    pb = api.PoaBatch(max_gpu_mem=MEM, max_sequences_per_poa=4)
    # empty batch: generate/get are no-ops
    pb.generate_poa()
    assert pb.get_consensus()[0] == [] and pb.get_total_poas() == 0
    # too long entry is skipped softly, too many entries too (cudapoa_batch.cuh:501-516)
    long_seq = b"A" * 1100
    st, per = pb.add_poa_group([(b"ACGTACGTAC", None), (b"ACGTACGTAC", None), (long_seq, None),
                                (b"ACGAACGTAC", None), (b"ACGTACGTAC", None), (b"ACGTACGTAC", None)])
    assert st == 0 and per == [0, 0, 2, 0, 0, 3]
    # negative weights are an argument error (cudapoa_batch.cuh:533-537 throws)
    st, _ = pb.add_poa_group([(b"ACGT", None), (b"ACGT", np.asarray([1, -1, 1, 1], dtype=np.int8)), (b"ACGT", None)])
    assert st == 15
    # identical reads: consensus == read (Test_CudapoaBatch.cu)
    st, _ = pb.add_poa_group([(b"ACGTTGCAACGT", None)] * 4)
    assert st == 0 and pb.get_total_poas() == 2
    pb.generate_poa()
    cons, cov, status = pb.get_consensus()
    assert status.tolist() == [0, 0] and cons[1] == b"ACGTTGCAACGT" and (cov[1] == 4).all()
    # reset then reuse
    pb.reset()
    assert pb.get_total_poas() == 0
    st, _ = pb.add_poa_group([(b"ACGTTGCAACGT", None)] * 3)
    pb.generate_poa()
    cons, _, status = pb.get_consensus()
    assert cons == [b"ACGTTGCAACGT"] and status[0] == 0
    pb.close()


def test_short_windows_and_partial_span_layers(oracle):
    """< 3 sequences: backbone + false (window.cpp:68-71).  Layers that do not span the window are aligned
    to the subgraph between their begin and end (window.cpp:96-103, graph.cpp:592-683) -- on the device."""
    from common import partial_span_windows
    two = [(b"ACGTACGTAC", None, 0, 0), (b"ACGTTCGTAC", None, 0, 9)]
    b = WindowBatch.from_lists([two])
    cons, clen, pol, status, _ = api.polish_windows(b, M, X, G, mem_per_batch=MEM)
    assert not pol[0] and api.consensus_list(cons, clen)[0] == b"ACGTACGTAC"
    pb = partial_span_windows()
    order = api.processing_order(pb)
    for trim in (False, True):
        oc, _, _ = oracle.polish(pb, order, M, X, G, tgs=trim, trim=trim, threads=16)
        for via_adapter in (False, True):
            cons, clen, pol, status, _ = api.polish_windows(pb, M, X, G, tgs=trim, trim=trim, mem_per_batch=MEM,
                                                            via_adapter=via_adapter)
            assert pol.all() and api.consensus_list(cons, clen) == oc
    # untrimmed consensus AND coverage through the batch API, banded too
    oc, ocov, _ = oracle.polish(pb, order, M, X, G, tgs=False, trim=False, threads=16)
    for banded in (False, True):
        gc, gcov, st = gpu_untrimmed(pb, banded=banded)
        assert (st == 0).all() and gc == oc
        assert all((a == c).all() for a, c in zip(gcov, ocov))


def test_batch_full_backpressure_and_multi_batch_concurrency(oracle):
    """exceeded_maximum_poas back-pressure (cudapoa_batch.cuh:122-125) and the End2End pattern of
    2 and 4 concurrent batches (Test_CudapoaBatchEnd2End.cu:39-91): results must not depend on how
    windows are split into batches."""
    b = synth_windows(300, 500, 12, 0.1, seed=47)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=True, trim=True, threads=16)
    base = None
    for batches, cap in ((1, 0), (2, 37), (4, 16)):
        cons, clen, pol, status, launches = api.polish_windows(b, M, X, G, mem_per_batch=3 << 30, batches_per_device=batches,
                                                               max_windows_per_round=cap)
        got = api.consensus_list(cons, clen)
        assert pol.all() and got == oc
        assert launches >= (1 if cap == 0 else 300 // cap)
    # a batch whose arena is tiny must accept a prefix and report it
    pb = api.PoaBatch(max_gpu_mem=MEM)
    cap = pb.info()["arena_capacity"]
    n, _ = pb.add_windows(b)
    assert n == 300 and cap > 0
    pb.close()


def test_ten_thousand_windows_are_deterministic_across_batch_splits():
    """BASELINE config 2 at full size (10k windows): a checksum of all consensus strings must not
    depend on the batch split, and windows that are duplicates must give identical consensus."""
    b = synth_windows(5000, 500, 32, 0.15, seed=49)
    # duplicate the set so that window i and i+5000 are identical inputs
    dup = WindowBatch(
        win_seq_off=np.concatenate([b.win_seq_off, b.win_seq_off[1:] + b.win_seq_off[-1]]),
        seq_off=np.concatenate([b.seq_off, b.seq_off[1:] + b.seq_off[-1]]),
        bases=np.concatenate([b.bases, b.bases]), weights=np.concatenate([b.weights, b.weights]),
        has_weights=np.concatenate([b.has_weights, b.has_weights]),
        begins=np.concatenate([b.begins, b.begins]), ends=np.concatenate([b.ends, b.ends]))
    c1, l1, p1, _, _ = api.polish_windows(dup, M, X, G, banded=True, mem_per_batch=24 << 30)
    c2, l2, p2, _, _ = api.polish_windows(dup, M, X, G, banded=True, mem_per_batch=24 << 30, batches_per_device=3,
                                          max_windows_per_round=777)
    assert p1.all() and p2.all()
    assert (l1 == l2).all() and (c1 == c2).all()
    assert (l1[:5000] == l1[5000:]).all() and (c1[:5000] == c1[5000:]).all()


def test_full_size_run_sampled_against_the_oracle(oracle):
    """BASELINE config 3 at full size: 10k windows, FULL band, every workspace reused by ~3 windows under full
    occupancy; a random sample of 192 windows must equal the oracle bit for bit, and the failure count is 0."""
    b = synth_windows(10000, 500, 32, 0.15, seed=53)
    cons, clen, pol, status, _ = api.polish_windows(b, M, X, G, banded=False, tgs=True, trim=True, mem_per_batch=48 << 30)
    assert pol.all() and (status == 0).all()
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(10000, size=192, replace=False))
    sub = WindowBatch.from_lists([[(s, w, bg, en) for s, w, bg, en in zip(*b.window(int(i)))] for i in pick])
    order = api.processing_order(sub)
    oc, _, _ = oracle.polish(sub, order, M, X, G, tgs=True, trim=True, threads=16)
    got = api.consensus_list(cons, clen)
    assert [got[int(i)] for i in pick] == oc


def test_long_window_stress_at_true_size(oracle):
    """BASELINE config 5: 1024 bp x 64 reads, 12 % error.  Half of the reads are longer than racon's hard-coded
    1023-base limit (cudabatch.cpp:59 BatchConfig(1023, ...)); the engine's limit is a BatchConfig field, so the
    stress case runs with max_sequence_size = 1279 (graph limits scale with it like batch.cu:34-71): full band
    bit-exact, static band within the stated tolerance."""
    b = synth_windows(24, 1024, 64, 0.12, seed=5)
    assert (np.diff(b.seq_off) > 1023).any()
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16)
    gc, gcov, st = gpu_untrimmed(b, banded=False, max_sequence_size=1279, mem=24 << 30)
    assert (st == 0).all() and gc == oc
    assert all((a == c).all() for a, c in zip(gcov, ocov))
    bc, _, st = gpu_untrimmed(b, banded=True, max_sequence_size=1279, mem=24 << 30)
    assert (st == 0).all()
    assert sum(a == c for a, c in zip(bc, oc)) >= 23 and max(edit_distance(a, c) for a, c in zip(bc, oc)) <= 2
    # with racon's own limit the long reads are dropped and reported per sequence, never silently
    pb = api.PoaBatch(max_gpu_mem=MEM, banded=True)
    n, seqs_added = pb.add_windows(b)
    pb.close()
    assert n == 24 and (np.asarray(seqs_added) < 64).any()


@pytest.mark.parametrize("scoring", [(3, -5, -4), (5, -4, -8), (1, -1, -1), (2, -3, -2)])
def test_awkward_windows_under_several_scoring_schemes(oracle, scoring):
    """Short and long layers, partial spans, qualities, up to 30 % error; full band: consensus and coverage are
    the oracle's, whatever the (match, mismatch, gap) triple (racon -m/-x/-g)."""
    from common import awkward_windows
    m, x, g = scoring
    b = awkward_windows(m, x, g)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, m, x, g, tgs=False, trim=False, threads=16)
    gc, gcov, st = gpu_untrimmed(b, banded=False, match=m, mismatch=x, gap=g)
    assert (st == 0).all() and gc == oc
    assert all((a == c).all() for a, c in zip(gcov, ocov))


def test_windows_built_in_the_columnar_arena_polish_like_the_oracle(oracle):
    """windows added through b200poa_arena_* (racon's createWindow/add_layer contract, layers of
    different windows interleaved) and polished with b200poa_polisher_polish_arena."""
    b = synth_windows(40, 300, 14, 0.12, seed=61, with_quality=True)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=True, trim=True, threads=16)
    per_win = [b.window(w) for w in range(b.n_windows)]
    qual = lambda w: bytes((np.asarray(w, dtype=np.int16) + 33).astype(np.uint8))
    arena = api.WindowArena()
    for s, w, _, _ in per_win:
        assert arena.add_window(bytes(s[0]), qual(w[0])) >= 0
    for i in range(1, 15):
        for wid, (s, w, bg, en) in enumerate(per_win):
            assert arena.add_layer(wid, bytes(s[i]), qual(w[i]), int(bg[i]), int(en[i]))
    built = arena.finalize()
    assert (built.bases == b.bases).all() and (built.weights == b.weights).all() and (built.begins == b.begins).all()
    pol = api.Polisher(devices=[0], batches_per_device=2, mem_per_batch=MEM)
    cons, clen, polished, status = pol.polish_arena(arena, tgs=True, trim=True, max_windows_per_round=16)
    pol.close()
    arena.close()
    assert polished.all() and api.consensus_list(cons, clen) == oc


def test_windows_the_batch_limits_cut_come_back_as_backbone_never_as_holes(oracle):
    """a window with < 3 sequences, or whose layers (or backbone) the batch limits dropped, is reported
    unpolished and holds its backbone (window.cpp:68-71 / cudapolisher.cpp:354-383 leave it to the CPU path); with
    accept_truncated the polisher does what the reference GPU adapter does (cudabatch.cpp:134-153, 232-233)."""
    deep = synth_windows(3, 200, 9, 0.1, seed=71)              # 10 sequences per window
    seqs, wts, bg, en = deep.window(0)
    two = [(seqs[0], wts[0], 0, 0), (seqs[1], None, 0, len(seqs[0]) - 1)]
    long_layer = [(s, w, b, e) for s, w, b, e in zip(*deep.window(1))]
    long_layer[3] = (long_layer[3][0] * 8, None, long_layer[3][2], long_layer[3][3])      # 1600 bases > 1023
    plain = [(s, w, b, e) for s, w, b, e in zip(*deep.window(2))]
    b = WindowBatch.from_lists([two, long_layer, plain])
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=True, trim=True, threads=4)
    bb = [b.window(w)[0][0] for w in range(3)]
    # exact mode (default): only the untouched window is polished
    pol = api.Polisher(devices=[0], mem_per_batch=MEM)
    cons, clen, polished, status = pol.polish(b, tgs=True, trim=True)
    pol.close()
    got = api.consensus_list(cons, clen)
    assert polished.tolist() == [False, False, True]
    assert got[0] == bb[0] and got[1] == bb[1] and got[2] == oc[2]
    assert status[1] == api.EXCEEDED_MAXIMUM_SEQUENCE_SIZE
    # depth limit 4 -> every deep window is cut; exact mode returns backbones, reference mode polishes the prefix
    pol = api.Polisher(devices=[0], mem_per_batch=MEM, max_sequences_per_poa=4)
    cons, clen, polished, status = pol.polish(b, tgs=True, trim=True)
    pol.close()
    assert polished.tolist() == [False, False, False] and api.consensus_list(cons, clen) == bb
    assert status[2] == api.EXCEEDED_MAXIMUM_SEQUENCES_PER_POA
    pol = api.Polisher(devices=[0], mem_per_batch=MEM, max_sequences_per_poa=4, accept_truncated=True)
    cons, clen, polished, status = pol.polish(b, tgs=True, trim=True)
    pol.close()
    assert polished.tolist() == [False, True, True]
    # the truncated result is the consensus of the first 4 sequences in processing order, trimmed at (4 - 1) / 2
    s0 = int(b.win_seq_off[2])
    keep = [int(i) for i in order[s0:s0 + 4]]
    sub = WindowBatch.from_lists([[plain[i] for i in keep]])
    tc, _, _ = oracle.polish(sub, identity_order(sub), M, X, G, tgs=True, trim=True, threads=1)
    assert api.consensus_list(cons, clen)[2] == tc[0]
    # the Python mirror of the adapter applies the same rule
    proc = api.CUDABatchProcessor(avail_mem=MEM)
    assert proc.add_windows(b) == 3
    out, ok = proc.generate_consensus()
    assert ok == [False, False, True] and out == [bb[0], bb[1], oc[2]]
    # a backbone longer than the limit: nothing is staged for the window, the status says why
    pb = api.PoaBatch(max_gpu_mem=MEM)
    st, per = pb.add_poa_group([(b"ACGT" * 300, None), (b"ACGTACGTAC", None), (b"ACGTACGTAC", None)])
    assert st == 0 and per == [2, 2, 2]
    pb.generate_poa()
    cons, _, status = pb.get_consensus()
    pb.close()
    assert status.tolist() == [api.EXCEEDED_MAXIMUM_SEQUENCE_SIZE] and cons == [b""]


def test_device_side_trim_and_compact_outputs_match_the_host_rule(oracle):
    """The trim span computed by the kernel (WindowOut::trim) is window.cpp:118-139's, the compact D2H moves exactly
    sum(len) elements, and sequences with constant weights ship no weight bytes."""
    b = synth_windows(64, 300, 12, 0.12, seed=73)             # backbone weight 0 ('!'), reads without quality
    pb = api.PoaBatch(max_gpu_mem=MEM)
    n, _ = pb.add_windows(b)
    pb.generate_poa()
    cons, cov, st, trim = pb.get_consensus(with_trim=True)
    info = pb.info()
    pb.close()
    assert (st == 0).all()
    for c, v, (tb, te) in zip(cons, cov, trim):
        want, chimeric = api.trim_consensus(c, v, 13)
        assert (c if tb >= te else c[tb:te + 1]) == want
    n_bases = int(b.seq_off[-1])
    assert info["h2d_bytes"] < n_bases + 64 * 13 * 40          # bases + tables, no weight bytes
    assert info["d2h_bytes"] <= 3 * sum((len(c) + 15) // 16 * 16 for c in cons) + 64 * 16 + 8
    q = synth_windows(16, 300, 12, 0.12, seed=73, with_quality=True)  # real qualities do travel
    pb = api.PoaBatch(max_gpu_mem=MEM)
    pb.add_windows(q)
    pb.generate_poa()
    pb.get_consensus()
    assert pb.info()["h2d_bytes"] > 1.8 * int(q.seq_off[-1]) - 16 * 320
    pb.close()


def test_single_process_multi_device_polisher(oracle):
    """racon's own model: ONE process, batch processors on every device (cudapolisher.cpp:228-240).  Needs >= 2 GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible device")
    b = synth_windows(600, 300, 10, 0.1, seed=79)
    order = api.processing_order(b)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=True, trim=True, threads=16)
    pol = api.Polisher(devices=list(range(n)), batches_per_device=2, mem_per_batch=MEM)
    cons, clen, polished, status = pol.polish(b, tgs=True, trim=True, max_windows_per_round=40)
    launches = pol.last["kernel_launches"]
    pol.close()
    assert polished.all() and api.consensus_list(cons, clen) == oc and launches >= 15


def test_adaptive_band_on_the_gpu(oracle):
    """Adaptive band (static band + retry with twice the width when the traceback nears the band edge): exact on the
    long-deletion windows the static band gets wrong, and bit-identical to its CPU emulation."""
    from emu_lib import Emu
    from test_emu import deletion_windows
    b = deletion_windows()
    order = identity_order(b)
    oc, ocov, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16, stride=8192)
    wins = [b.window(w)[:2] for w in range(b.n_windows)]
    out = {}
    for mode in (True, "adaptive"):
        pb = api.PoaBatch(max_gpu_mem=MEM, banded=mode)
        for seqs, wts in wins:  # the backbone carries its weight-0 quality, the reads none: same input as the oracle's
            assert pb.add_poa_group(list(zip(seqs, wts)))[0] == 0
        pb.generate_poa()
        out[mode] = pb.get_consensus()
        pb.close()
    gc, gcov, st = out["adaptive"]
    assert (st == 0).all() and gc == oc and all((a == c).all() for a, c in zip(gcov, ocov))
    sc, _, sst = out[True]
    ec, _, est, _ = Emu().polish(b, order, M, X, G, max_nodes=4092, max_edges=24000, band=256, stride=8192)
    assert (sst == est).all() and sc == ec and sc != oc
    # on ordinary windows the adaptive band is the static band (no retry): same result, same tolerance
    a = synth_windows(64, 500, 32, 0.15, seed=83)
    r_static = gpu_untrimmed(a, banded=True)
    r_adapt = gpu_untrimmed(a, banded="adaptive")
    assert r_static[0] == r_adapt[0]


def test_int32_cells_on_the_gpu(oracle, monkeypatch):
    """32-bit score cells (spoa's int32 switch): a scoring scheme int16 cannot hold runs through them on its own, and
    forced on ordinary windows (B200POA_FORCE_CELLS32, read at batch creation) they are bit-exact too."""
    b = synth_windows(8, 500, 32, 0.15, seed=3)
    order = api.processing_order(b)
    oc, ocov, _ = oracle.polish(b, order, 120, -120, -120, tgs=False, trim=False, threads=16)
    gc, gcov, st = gpu_untrimmed(b, match=120, mismatch=-120, gap=-120)
    assert (st == 0).all() and gc == oc and all((a == c).all() for a, c in zip(gcov, ocov))
    monkeypatch.setenv("B200POA_FORCE_CELLS32", "1")
    from common import partial_span_windows
    for w in (synth_windows(24, 300, 16, 0.15, seed=9, with_quality=True), partial_span_windows()):
        order = api.processing_order(w)
        oc, ocov, _ = oracle.polish(w, order, M, X, G, tgs=False, trim=False, threads=16)
        for banded in (False, True):
            gc, gcov, st = gpu_untrimmed(w, banded=banded)
            assert (st == 0).all() and gc == oc and all((a == c).all() for a, c in zip(gcov, ocov))
