"""Regenerates tests/golden/lambda_windows.npz and tests/golden/cudapoa_windows.npz (run in the container that
has /root/reference; needs `make -C oracle dump`).

lambda_windows.npz -- REAL racon windows.  oracle/_ref/racon_dump (the unmodified racon CPU pipeline compiled in
place + oracle/racon_dump.cpp) runs Polisher::initialize on the reference's own test data
(/root/reference/test/data/sample_{reads.fastq,reads.fasta,overlaps.paf,layout.fasta,reference.fasta}.gz) with the
parameters of test/racon_test.cpp:88-130,176-196 (kC, quality 10, error 0.3, 5/-4/-8, trim) and records every
window (layers, qualities, spans, type) together with the consensus racon's CPU path computes for it; the dumper
itself checks the stitched contig against the reference's golden edit distances (1312 / 1566 / 1289).
    <case>_bases      uint8  2-bit packed ACGT codes (all sample bases are ACGT; asserted)
    <case>_qual       uint8  PHRED+33 characters of the sequences that have a quality string, concatenated
    <case>_seq_len / _begin / _end / _has_q   per sequence, ADD order, sequence 0 of a window = backbone
    <case>_win_nseq   sequences per window;  <case>_cons / _cons_len / _polished  racon CPU result per window
    <case>_params     [window_length, tgs, trim, m, x, g, edit_distance, contig_length]
    reference         2-bit packed sample_reference.fasta (the target of the edit-distance goldens), reference_len

cudapoa_windows.npz -- the 67 deep windows of vendor/GenomeWorks/cudapoa/data/sample-windows.txt (depth 105-170,
format: count line, then one read per line; cudapoa/include/.../utils.hpp:97-139) and the consensus + coverage the
unmodified reference (oracle/_ref: spoa called like window.cpp:73-116, all reads full-span, weight 1, 3/-5/-4)
produces for them in file order.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
DATA = os.path.join(REF, "test/data")
CODE = np.full(256, 255, dtype=np.uint8)
for i, c in enumerate(b"ACGT"):
    CODE[c] = i


def pack2(b: np.ndarray) -> np.ndarray:
    c = CODE[b]
    assert (c < 4).all(), "non-ACGT base in the sample data"
    pad = (-c.shape[0]) % 4
    c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def parse_dump(path):
    raw = open(path, "rb").read()
    pos = 0

    def u32():
        nonlocal pos
        v = int.from_bytes(raw[pos:pos + 4], "little")
        pos += 4
        return v

    assert u32() == 0x31445752
    n_win, tgs, trim = u32(), u32(), u32()
    m, x, g = [v - (1 << 32) if v >= (1 << 31) else v for v in (u32(), u32(), u32())]
    wl, ed, clen = u32(), u32(), u32()
    bases, quals, seq_len, begin, end, has_q, win_nseq, cons, cons_len, polished = [], [], [], [], [], [], [], [], [], []
    for _ in range(n_win):
        pos += 8  # target id
        pos += 4  # rank
        ns = u32()
        polished.append(raw[pos]); pos += 1
        cl = u32()
        cons.append(raw[pos:pos + cl]); pos += cl
        cons_len.append(cl)
        win_nseq.append(ns)
        for _ in range(ns):
            ln, bg, en = u32(), u32(), u32()
            hq = raw[pos]; pos += 1
            bases.append(raw[pos:pos + ln]); pos += ln
            if hq:
                quals.append(raw[pos:pos + ln]); pos += ln
            seq_len.append(ln); begin.append(bg); end.append(en); has_q.append(hq)
    assert pos == len(raw)
    return {
        "bases": pack2(np.frombuffer(b"".join(bases), dtype=np.uint8)),
        "qual": np.frombuffer(b"".join(quals), dtype=np.uint8),
        "seq_len": np.asarray(seq_len, dtype=np.int32), "begin": np.asarray(begin, dtype=np.int32),
        "end": np.asarray(end, dtype=np.int32), "has_q": np.asarray(has_q, dtype=np.uint8),
        "win_nseq": np.asarray(win_nseq, dtype=np.int32), "cons": np.frombuffer(b"".join(cons), dtype=np.uint8),
        "cons_len": np.asarray(cons_len, dtype=np.int32), "polished": np.asarray(polished, dtype=np.uint8),
        "params": np.asarray([wl, tgs, trim, m, x, g, ed, clen], dtype=np.int64),
    }


def lambda_windows():
    import gzip
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dump"], check=True)
    dump = os.path.join(ROOT, "oracle", "_ref", "racon_dump")
    out = {}
    expect = {"fastq_500": 1312, "fasta_500": 1566, "fastq_1000": 1289}  # test/racon_test.cpp:106,128,194
    for case in expect:
        kind, wl = case.split("_")
        tmp = f"/tmp/lambda_{case}.bin"
        subprocess.run([dump, f"{DATA}/sample_reads.{kind}.gz", f"{DATA}/sample_overlaps.paf.gz",
                        f"{DATA}/sample_layout.fasta.gz", f"{DATA}/sample_reference.fasta.gz", wl, "10", "0.3", "5", "-4",
                        "-8", tmp], check=True, stderr=subprocess.DEVNULL)
        d = parse_dump(tmp)
        assert int(d["params"][6]) == expect[case], (case, d["params"])
        for k, v in d.items():
            out[f"{case}_{k}"] = v
    ref = b"".join(l.strip() for l in gzip.open(f"{DATA}/sample_reference.fasta.gz", "rb").read().split(b"\n")[1:])
    out["reference"] = pack2(np.frombuffer(ref.upper(), dtype=np.uint8))
    out["reference_len"] = np.asarray([len(ref)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "lambda_windows.npz"), **out)


def cudapoa_windows():
    from oracle_lib import Ref
    r = Ref()
    assert r.available
    lines = open(os.path.join(REF, "vendor/GenomeWorks/cudapoa/data/sample-windows.txt"), "rb").read().split(b"\n")
    wins, i = [], 0
    while i < len(lines) and lines[i].strip():
        n = int(lines[i]); i += 1
        wins.append([l.strip() for l in lines[i:i + n]]); i += n
    assert len(wins) == 67
    cons, cov = [], []
    for w in wins:
        c, v, _ = r.spoa_window(w, [None] * len(w), 3, -5, -4)
        cons.append(c); cov.append(v.astype(np.uint16))
    flat = b"".join(s for w in wins for s in w)
    np.savez_compressed(os.path.join(HERE, "cudapoa_windows.npz"),
                        bases=pack2(np.frombuffer(flat, dtype=np.uint8)),
                        seq_len=np.asarray([len(s) for w in wins for s in w], dtype=np.int32),
                        win_nseq=np.asarray([len(w) for w in wins], dtype=np.int32),
                        cons=np.frombuffer(b"".join(cons), dtype=np.uint8),
                        cons_len=np.asarray([len(c) for c in cons], dtype=np.int32), cov=np.concatenate(cov))


if __name__ == "__main__":
    lambda_windows()
    cudapoa_windows()
    for f in ("lambda_windows.npz", "cudapoa_windows.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
