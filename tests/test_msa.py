"""Multiple sequence alignment output (Batch::get_msa, generateMSAKernel).

Parity target = spoa::Graph::generate_multiple_sequence_alignment (vendor/spoa/src/graph.cpp:373-427), which is what the
reference's own test holds cudapoa's MSA against (vendor/GenomeWorks/cudapoa/tests/Test_CudapoaGenerateMSA2.cu:117-128).
CPU here: the oracle's restatement against committed digests of the unmodified reference (tests/golden/ref_msa.npz),
live against oracle/_ref, and the engine's emulation against the oracle.  GPU: tests/test_gpu_msa.py."""
import os
import sys

import numpy as np
import pytest

from common import G, GOLDEN, M, X
from oracle_lib import oracle_window_msa, ref_window_msa

sys.path.insert(0, GOLDEN)
from make_msa_golden import digest, msa_groups  # noqa: E402


@pytest.fixture(scope="module")
def groups():
    return msa_groups()


def test_oracle_msa_matches_the_committed_reference_digests(oracle, groups):
    z = np.load(os.path.join(GOLDEN, "ref_msa.npz"))
    assert set(z.files) == set(groups)
    for name, (seqs, wts, (m, x, g)) in groups.items():
        n_rows, msa_len, sha = z[name].tobytes().decode().split()
        rows = oracle_window_msa(oracle, seqs, wts, None, None, m, x, g)
        assert (len(rows), len(rows[0])) == (int(n_rows), int(msa_len)), name
        assert digest(rows) == sha, name
        # Test_CudapoaGenerateMSA2.cu:120-126: a row without its gaps is the input sequence
        assert [r.replace(b"-", b"") for r in rows] == [bytes(s) for s in seqs]


def test_oracle_msa_matches_live_reference(oracle, ref, groups):
    if not ref.available:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for name, (seqs, wts, (m, x, g)) in list(groups.items())[:12]:
        assert oracle_window_msa(oracle, seqs, wts, None, None, m, x, g) == ref_window_msa(ref, seqs, wts, m, x, g), name
