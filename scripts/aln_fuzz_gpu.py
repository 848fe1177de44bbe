"""GPU fuzz of the overlap aligner through the C ABI against the unmodified edlib (oracle/_ref must have travelled):
mixed shapes in one batch (short / tall / huge, unrelated pairs, truncated queries), several band guesses and memory
budgets (thin and saturated levels, one-warp and team launches, side stream).  usage: python scripts/aln_fuzz_gpu.py [seconds]"""
import sys, time, json
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from oracle_lib import Ref, ref_align
from racon_gpu_b200.aligner import CUDABatchAligner, pack_pairs

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutate(rng, t, err):
    keep = rng.random(len(t))
    out = []
    for c, x in zip(t, keep):
        if x < err / 3: out.append(ACGT[rng.integers(4)])
        elif x < 2 * err / 3: continue
        elif x < err: out.append(c); out.append(ACGT[rng.integers(4)])
        else: out.append(c)
    return bytes(out) if out else b"A"


def make_pairs(rng, count):
    pairs = []
    for _ in range(count):
        shape = rng.integers(0, 8)
        m = int({0: rng.integers(1, 400), 1: rng.integers(400, 4000), 2: rng.integers(4000, 12000), 3: rng.integers(12000, 26000),
                 4: rng.integers(1, 3000), 5: rng.integers(2000, 9000), 6: rng.integers(2000, 9000), 7: rng.integers(400, 4000)}[int(shape)])
        t = rng.choice(ACGT, size=m)
        if shape == 4:
            q = rng.choice(ACGT, size=int(rng.integers(1, 12000))).tobytes()
        else:
            q = mutate(rng, t, float(rng.choice([0.0, 0.01, 0.05, 0.12, 0.2, 0.35])))
            if rng.random() < 0.1: q = q[:max(1, len(q) // int(rng.integers(2, 6)))]
            if rng.random() < 0.05: q = q + rng.choice(ACGT, size=int(rng.integers(1, 3000))).tobytes()
        pairs.append((q, t.tobytes()))
    return pairs


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    r = Ref()
    assert r.available, "oracle/_ref missing"
    rng = np.random.default_rng(2026)
    t0, rounds, checked, bad = time.time(), 0, 0, 0
    while time.time() - t0 < seconds:
        count = int(rng.choice([40, 400, 3000]))
        pairs = make_pairs(rng, count)
        if count == 3000:  # saturate: many copies of the tall ones
            pairs = pairs + [p for p in pairs if len(p[0]) > 2048] * 3
        with ThreadPoolExecutor(16) as ex:
            want = list(ex.map(lambda p: ref_align(r, p[0], p[1]), pairs))
        q, qo, t, to = pack_pairs(pairs)
        for guess, mem in ((-1, 16 << 30), (int(rng.integers(1, 400)), 16 << 30), (0, 1 << 30)):
            al = CUDABatchAligner(device_id=0, max_gpu_memory=mem)
            al.set_band_guess(guess)
            for rnd in range(2 if guess == -1 else 1):
                first = 0
                while first < len(pairs):
                    n = al.add_overlaps(q, qo, t, to, first)
                    al.align_all()
                    text, off, ln, ed = al.cigars()
                    for k in range(n):
                        ops, score, cig = want[first + k]
                        if ed[k] != score or text[off[k]:off[k] + ln[k]] != cig:
                            bad += 1
                            print("MISMATCH", len(pairs[first + k][0]), len(pairs[first + k][1]), guess, mem, flush=True)
                        checked += 1
                    al.reset()
                    first += n
            al.close()
        rounds += 1
    print(json.dumps({"rounds": rounds, "alignments_checked": checked, "mismatches": bad, "seconds": round(time.time() - t0)}))


if __name__ == "__main__":
    main()
