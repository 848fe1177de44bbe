"""Regenerates tests/golden/ref_msa.npz (run in the container that has /root/reference).

Multiple sequence alignments of the UNMODIFIED reference -- spoa::Graph::generate_multiple_sequence_alignment driven
exactly like the reference's own MSA test drives it (vendor/GenomeWorks/cudapoa/tests/Test_CudapoaGenerateMSA2.cu:62-79)
through oracle/_ref (oracle/ref_driver.cpp: ref_spoa_window_msa) -- for groups the GPU box can rebuild from the other
committed fixtures: spoa's sample.fastq group, seeded synthetic windows, the first deep cudapoa sample windows.
Stored per group: number of rows, MSA length, SHA-256 of the rows joined by '\n' (the alignments themselves are a few
MB; the inputs are already in the repository).
"""
import hashlib, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def msa_groups():
    """name -> (list of sequences, list of weights or None, (m, x, g)); sequences in the order they are added."""
    from common import cudapoa_fixture, spoa_window
    from racon_gpu_b200.windows import synth_windows
    groups = {}
    for q in (False, True):
        b = spoa_window(q)
        seqs, wts, _, _ = b.window(0)
        groups["spoa_sample_q%d" % int(q)] = (seqs, wts, (5, -4, -8))
    for name, (n, L, D, e, wq, seed) in {"A": (6, 500, 32, 0.15, False, 31), "Q": (6, 300, 14, 0.2, True, 32),
                                         "S": (8, 60, 5, 0.25, False, 33)}.items():
        b = synth_windows(n, L, D, e, seed=seed, with_quality=wq)
        for w in range(b.n_windows):
            seqs, wts, _, _ = b.window(w)
            groups["synth_%s_%d" % (name, w)] = (seqs, wts, (3, -5, -4))
    b, _, _ = cudapoa_fixture()
    for w in range(4):
        seqs, wts, _, _ = b.window(w)
        groups["cudapoa_%d" % w] = (seqs, wts, (3, -5, -4))
    return groups


def digest(rows):
    return hashlib.sha256(b"\n".join(rows)).hexdigest()


if __name__ == "__main__":
    from oracle_lib import Ref, ref_window_msa
    r = Ref(); assert r.available
    out = {}
    for name, (seqs, wts, (m, x, g)) in msa_groups().items():
        rows = ref_window_msa(r, seqs, wts, m, x, g)
        assert [row.replace(b"-", b"") for row in rows] == [bytes(s) for s in seqs]
        out[name] = np.frombuffer(("%d %d %s" % (len(rows), len(rows[0]), digest(rows))).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_msa.npz"), **out)
    print("wrote", len(out), "MSA digests")
