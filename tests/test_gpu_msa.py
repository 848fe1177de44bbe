"""GPU (-m gpu): multiple sequence alignment output (Batch::get_msa) through the C ABI.

Parity target = spoa::Graph::generate_multiple_sequence_alignment (vendor/spoa/src/graph.cpp:373-427), the value the
reference's own test expects from cudapoa (cudapoa/tests/Test_CudapoaGenerateMSA2.cu:85-130): committed digests of the
unmodified reference (tests/golden/ref_msa.npz) and the oracle's restatement on everything else."""
import os
import sys

import numpy as np
import pytest

from common import G, GOLDEN, M, X, cudapoa_fixture, lambda_fixture, partial_span_windows
from oracle_lib import oracle_window_msa, window_sequences
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows

sys.path.insert(0, GOLDEN)
from make_msa_golden import digest, msa_groups  # noqa: E402

pytestmark = pytest.mark.gpu
MEM = 6 << 30
BOTH = api.OUTPUT_CONSENSUS | api.OUTPUT_MSA


def test_msa_of_groups_equals_the_reference_digests():
    """Test_CudapoaGenerateMSA2.cu:85-130 with the reference's expected value taken from the committed digests."""
    z = np.load(os.path.join(GOLDEN, "ref_msa.npz"))
    groups = msa_groups()
    by_scoring = {}
    for name, (seqs, wts, sc) in groups.items():
        by_scoring.setdefault(sc, []).append((name, seqs, wts))
    for (m, x, g), items in by_scoring.items():
        pb = api.PoaBatch(max_gpu_mem=MEM, gap=g, mismatch=x, match=m, output_mask=api.OUTPUT_MSA)
        for _, seqs, wts in items:
            st, per = pb.add_poa_group([(s, w) for s, w in zip(seqs, wts)])
            assert st == 0 and all(p == 0 for p in per)
        pb.generate_poa()
        msa, status = pb.get_msa()
        with pytest.raises(RuntimeError):  # cudapoa_batch.cuh:205-209: consensus was not asked for
            pb.get_consensus()
        pb.close()
        assert (status == 0).all()
        for (name, seqs, _), rows in zip(items, msa):
            n_rows, n_cols, sha = z[name].tobytes().decode().split()
            assert (len(rows), len(rows[0])) == (int(n_rows), int(n_cols)), name
            assert digest(rows) == sha, name
            assert [r.replace(b"-", b"") for r in rows] == [bytes(s) for s in seqs]  # :120-126


def _columnar_msa(b, m, x, g, banded=False, mask=BOTH, **kw):
    pb = api.PoaBatch(max_gpu_mem=MEM, gap=g, mismatch=x, match=m, banded=banded, output_mask=mask, **kw)
    first, msas, stats, cons = 0, [], [], []
    rounds = 0
    while first < b.n_windows:  # a full MSA arena is back-pressure ("exceeded_maximum_poas"): launch and go on
        n, _ = pb.add_windows(b, first)
        assert n > 0
        pb.generate_poa()
        ms, st = pb.get_msa()
        msas += ms
        stats += st.tolist()
        if mask & api.OUTPUT_CONSENSUS:
            cons += pb.get_consensus()[0]
        pb.reset()
        first += n
        rounds += 1
    pb.close()
    return msas, np.asarray(stats), cons, rounds


def _oracle_msas(oracle, b, m, x, g):
    order = api.processing_order(b)
    out = []
    for w in range(b.n_windows):
        seqs, wts, bg, en = window_sequences(b, order, w)
        out.append(oracle_window_msa(oracle, seqs, wts, bg, en, m, x, g))
    return out, order


def test_msa_and_consensus_together_on_synthetic_and_partial_span_windows(oracle):
    for b in (synth_windows(48, 500, 32, 0.15, seed=51), synth_windows(24, 300, 14, 0.2, seed=52, with_quality=True),
              partial_span_windows()):
        want, order = _oracle_msas(oracle, b, M, X, G)
        oc, _, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False, threads=16)
        msas, st, cons, _ = _columnar_msa(b, M, X, G)
        assert (st == 0).all()
        assert msas == want
        assert cons == oc  # asking for the MSA changes nothing in the consensus


def test_msa_of_real_racon_windows(oracle):
    """Real lambda-phage windows: partial-span layers (aligned to subgraphs), FASTQ qualities, uneven depth."""
    b, _, _, p = lambda_fixture("fastq_500")
    want, _ = _oracle_msas(oracle, b, p["m"], p["x"], p["g"])
    msas, st, _, _ = _columnar_msa(b, p["m"], p["x"], p["g"], mask=api.OUTPUT_MSA)
    nseq = np.diff(b.win_seq_off)
    for w in range(b.n_windows):
        if st[w] == 0:
            assert msas[w] == want[w], w
        else:  # racon_gpu's limit, mirrored: alignments of max_consensus_size (2046) columns or more
            assert st[w] == 2 and len(want[w][0]) >= 2046
    assert (st == 0).sum() >= 0.9 * b.n_windows and nseq.max() > 3


def test_msa_of_deep_cudapoa_windows(oracle):
    """sample-windows.txt (depth 105-170, the reference's MSA benchmark input): rows equal spoa's."""
    b, _, _ = cudapoa_fixture()
    wins = [b.window(w)[0] for w in range(20)]
    want = [oracle_window_msa(oracle, seqs, [None] * len(seqs), None, None, M, X, G) for seqs in wins]
    pb = api.PoaBatch(max_gpu_mem=MEM, output_mask=api.OUTPUT_MSA)
    for seqs in wins:
        st, per = pb.add_poa_group([(s, None) for s in seqs])
        assert st == 0 and all(x == 0 for x in per)
    pb.generate_poa()
    msa, status = pb.get_msa()
    pb.close()
    assert (status == 0).all() and msa == want


def test_msa_wider_than_max_consensus_size_is_reported_like_cudapoa(oracle):
    """cudapoa_generate_msa.cuh:203-208 / Test_CudapoaGenerateMSA2.cu:132-170 (CudapoaMSAFailure): an alignment of
    max_consensus_size (2 x max_sequence_size) columns or more is exceeded_maximum_sequence_size; the window's
    consensus and the other windows are not affected."""
    b = synth_windows(6, 240, 16, 0.3, seed=55)
    want, order = _oracle_msas(oracle, b, M, X, G)
    oc, _, _ = oracle.polish(b, order, M, X, G, tgs=False, trim=False)
    msas, st, cons, _ = _columnar_msa(b, M, X, G, max_sequence_size=280)
    wide = [len(r[0]) >= 560 for r in want]
    assert any(wide) and not all(wide)
    for w in range(b.n_windows):
        assert (st[w], msas[w]) == ((2, None) if wide[w] else (0, want[w]))
    assert cons == oc


def test_banded_msa_rows_spell_their_sequences():
    """Static band (no exact oracle): every row without its gaps is its input sequence, all rows of a window have one
    length, no column is empty -- at BASELINE config A's shape, 2000 windows, through several rounds of a small arena."""
    b = synth_windows(2000, 500, 32, 0.15, seed=53)
    order = api.processing_order(b)
    os.environ["B200POA_MSA_ARENA_MB"] = "40"
    try:
        msas, st, _, rounds = _columnar_msa(b, M, X, G, banded=True, mask=api.OUTPUT_MSA)
    finally:
        del os.environ["B200POA_MSA_ARENA_MB"]
    assert rounds > 1 and (st == 0).all() and len(msas) == b.n_windows
    for w in range(0, b.n_windows, 7):
        seqs, _, _, _ = window_sequences(b, order, w)
        rows = msas[w]
        assert [r.replace(b"-", b"") for r in rows] == seqs
        mat = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), -1)
        assert (mat != ord("-")).any(axis=0).all()


def test_get_msa_is_unavailable_without_the_output_mask():
    pb = api.PoaBatch(max_gpu_mem=MEM)
    b = synth_windows(2, 100, 4, 0.1, seed=54)
    pb.add_windows(b)
    pb.generate_poa()
    with pytest.raises(RuntimeError, match="output_type_unavailable"):
        pb.get_msa()
    pb.close()
