"""Regenerates tests/golden/lambda_overlaps.npz (run in the container that has /root/reference).

REAL overlap alignment inputs: the 181 read-to-contig overlaps of the reference's own test data
(/root/reference/test/data/sample_overlaps.paf.gz over sample_reads.fasta.gz and sample_layout.fasta.gz), cut out
exactly like racon cuts them before aligning (src/overlap.cpp:186-199: the read segment [q_begin, q_end), reverse
complemented for '-' overlaps; the contig segment [t_begin, t_end)), each aligned by the UNMODIFIED edlib the way racon's
CPU path calls it (oracle/_ref: ref_edlib_nw = src/overlap.cpp:205-224).
    q_bases / t_bases   uint8  2-bit packed ACGT, all query / target segments back to back
    q_len / t_len       int32  per overlap
    score               int32  edit distance
    cigar_sha           bytes  per overlap "n_ops sha256(cigar)\n" (the CIGAR strings total 2.5 MB; the inputs are here)
    q_first / t_begin   int32  where the segments start in read (strand-adjusted, src/overlap.cpp:241) / contig coordinates
    bp / bp_count       uint32 (t, q) breaking points of the UNMODIFIED racon::Overlap::find_breaking_points fed edlib's
                               CIGAR, window length 500 (oracle/_ref: ref_racon_breaking_points = src/overlap.cpp:226-290)
"""
import gzip, hashlib, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
DATA = "/root/reference/test/data"
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def fasta(path):
    out, name = {}, None
    for line in gzip.open(path, "rt"):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            out[name] = []
        elif name is not None:
            out[name].append(line)
    return {k: "".join(v).upper().encode() for k, v in out.items()}


if __name__ == "__main__":
    from make_lambda_golden import pack2
    from oracle_lib import Ref, ref_align, ref_breaking_points
    r = Ref(); assert r.available
    reads, contigs = fasta(os.path.join(DATA, "sample_reads.fasta.gz")), fasta(os.path.join(DATA, "sample_layout.fasta.gz"))
    qs, ts, scores, shas, q_first, t_begin, bps, bp_count = [], [], [], [], [], [], [], []
    for line in gzip.open(os.path.join(DATA, "sample_overlaps.paf.gz"), "rt"):
        f = line.split("\t")
        qn, ql, qb, qe, strand, tn, tl, tb, te = f[0], int(f[1]), int(f[2]), int(f[3]), f[4], f[5], int(f[6]), int(f[7]), int(f[8])
        read = reads[qn]
        assert len(read) == ql and len(contigs[tn]) == tl
        q = read[qb:qe] if strand == "+" else read.translate(COMP)[::-1][ql - qe:ql - qe + (qe - qb)]
        t = contigs[tn][tb:te]
        ops, score, cigar = ref_align(r, q, t)
        qs.append(q); ts.append(t); scores.append(score)
        shas.append(("%d %s\n" % (ops.shape[0], hashlib.sha256(cigar).hexdigest())).encode())
        bp = ref_breaking_points(r, cigar, ql, qb, qe, 1 if strand == "-" else 0, tl, tb, te, 500)
        q_first.append(ql - qe if strand == "-" else qb); t_begin.append(tb); bps.append(bp.reshape(-1)); bp_count.append(bp.shape[0])
    np.savez_compressed(os.path.join(HERE, "lambda_overlaps.npz"),
                        q_bases=pack2(np.frombuffer(b"".join(qs), dtype=np.uint8)),
                        t_bases=pack2(np.frombuffer(b"".join(ts), dtype=np.uint8)),
                        q_len=np.asarray([len(x) for x in qs], dtype=np.int32),
                        t_len=np.asarray([len(x) for x in ts], dtype=np.int32),
                        score=np.asarray(scores, dtype=np.int32), cigar_sha=np.frombuffer(b"".join(shas), dtype=np.uint8),
                        q_first=np.asarray(q_first, dtype=np.int32), t_begin=np.asarray(t_begin, dtype=np.int32),
                        bp=np.concatenate(bps).astype(np.uint32), bp_count=np.asarray(bp_count, dtype=np.int32))
    print("wrote", len(qs), "overlaps; mean identity-ish", 1 - sum(scores) / sum(len(x) for x in ts))
