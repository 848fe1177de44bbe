"""Kernel rate on REAL racon windows: the lambda-phage fixture (tests/golden/lambda_windows.npz), replicated to fill the
grid several times over.  Prints launch times, windows/s, failures and (B200POA_PHASE_TIMERS=1) the phase split."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from racon_gpu_b200 import api
from racon_gpu_b200.windows import WindowBatch
from common import lambda_fixture

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="fastq_500")
ap.add_argument("--copies", type=int, default=80)
ap.add_argument("--band", default="0")          # 0 full, 1 static 256, adaptive
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--mem-gb", type=float, default=64.0)
args = ap.parse_args()
b, cons, polished, prm = lambda_fixture(args.case)
wins = []
for w in range(b.n_windows):
    seqs, wts, bg, en = b.window(w)
    if len(seqs) < 3:
        continue
    wins.append([(seqs[i], wts[i], int(bg[i]), int(en[i])) for i in range(len(seqs))])
depth = np.mean([len(w) for w in wins])
big = WindowBatch.from_lists(wins * args.copies)
banded = {"0": False, "1": True}.get(args.band, args.band)
max_seq = 1023 if prm["window_length"] <= 500 else 2047
pb = api.PoaBatch(max_gpu_mem=int(args.mem_gb * (1 << 30)), banded=banded, max_sequence_size=max_seq,
                  match=prm["m"], mismatch=prm["x"], gap=prm["g"])
n, _ = pb.add_windows(big)
pb.upload()
torch.cuda.synchronize()
ms = []
for _ in range(args.launches):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pb.launch(); c.record(); torch.cuda.synchronize()
    ms.append(round(a.elapsed_time(c), 2))
pb.download()
_, _, st = pb.get_consensus()
print(f"case {args.case} band {args.band}: {len(wins)} real windows (mean depth {depth:.1f}) x {args.copies} = {n} staged of {big.n_windows};"
      f" launch ms {ms}; windows/s {round(n / (min(ms) / 1e3))}; failed {int((st != 0).sum())}")
pc = pb.phase_cycles()
if pc:
    tot = sum(pc.values())
    print("  phase share:", {k: round(100.0 * v / tot, 1) for k, v in pc.items()})
