"""Regenerates the golden fixtures under tests/golden/ (run in the container that has /root/reference).

1. spoa_sample.fastq.gz   -- the 55 reads of /root/reference/vendor/spoa/test/data/sample.fastq, the
   input of spoa's own known-answer tests (vendor/spoa/test/spoa_test.cpp:220-238 GlobalConsensus and
   :283-301 GlobalConsensusWithQualities: NW, m=5 x=-4 g=-8).  The expected strings are transcribed
   in spoa_golden.json from those two tests.
2. ref_windows.npz        -- seeded synthetic windows + the consensus the UNMODIFIED reference
   (oracle/_ref: racon::Window::generate_consensus + spoa SIMD engine) produces for them, so the
   GPU box (which has no /root/reference) can still check against the real reference's outputs.
"""
import gzip, json, os, re, shutil, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

def spoa_golden():
    src = os.path.join(REF, "vendor/spoa/test/data/sample.fastq")
    with open(src, "rb") as fi, open(os.path.join(HERE, "spoa_sample.fastq.gz"), "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fo:
        shutil.copyfileobj(fi, fo)
    text = open(os.path.join(REF, "vendor/spoa/test/spoa_test.cpp")).read()
    out = {}
    for name in ("GlobalConsensus", "GlobalConsensusWithQualities"):
        body = text[text.index("TEST_F(SpoaAlignmentTest, %s)" % name):]
        body = body[:body.index("EXPECT_TRUE")]
        lit = body[body.index("valid_result ="):]
        out[name] = "".join(re.findall(r'"([ACGT]+)"', lit))
    out["scoring"] = {"m": 5, "x": -4, "g": -8}
    out["source"] = "vendor/spoa/test/spoa_test.cpp:220-238,283-301"
    json.dump(out, open(os.path.join(HERE, "spoa_golden.json"), "w"), indent=1)

def ref_windows():
    from racon_gpu_b200.windows import synth_windows
    from oracle_lib import Ref
    r = Ref(); assert r.available
    fixtures = {}
    cases = {"A": (24, 500, 32, 0.15, False), "C": (24, 500, 8, 0.05, False), "Q": (16, 400, 20, 0.12, True),
             "B": (4, 900, 64, 0.12, False)}
    for name, (n, L, D, e, q) in cases.items():
        b = synth_windows(n, L, D, e, seed=20260921, with_quality=q)
        for tgs_trim in (0, 1):
            cons, pol = r.polish(b, 3, -5, -4, tgs=bool(tgs_trim), trim=bool(tgs_trim), threads=8, window_length=L)
            fixtures["%s_cons_%d" % (name, tgs_trim)] = np.frombuffer(b"\n".join(cons), dtype=np.uint8)
        fixtures[name + "_params"] = np.asarray([n, L, D, int(e * 1000), int(q), 20260921], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_windows.npz"), **fixtures)

if __name__ == "__main__":
    spoa_golden(); ref_windows(); print("golden fixtures written to", HERE)
