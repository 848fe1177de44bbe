"""First-light check on a B200: GPU engine vs oracle on a few windows, plus a rough timing."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from racon_gpu_b200.windows import synth_windows
from racon_gpu_b200 import api
from oracle_lib import Oracle

def run(tag, nwin, L, D, e, banded, q=False, mem=16 << 30):
    b = synth_windows(nwin, L, D, e, seed=21, with_quality=q)
    order = api.processing_order(b)
    o = Oracle()
    ncheck = min(nwin, 64)
    oc, ocov, _ = o.polish(b.slice(0, ncheck), order[:int(b.win_seq_off[ncheck])], 3, -5, -4, tgs=False, trim=False, threads=16)
    pb = api.PoaBatch(max_gpu_mem=mem, banded=banded)
    print(tag, "info", pb.info(), flush=True)
    n, _ = pb.add_windows(b)
    assert n == nwin, (n, nwin)
    torch.cuda.synchronize()
    t = time.time(); pb.generate_poa(); gc, gcov, st = pb.get_consensus(); dt = time.time() - t
    same = sum(a == c for a, c in zip(oc, gc[:ncheck]))
    csame = sum(len(a) == len(c) and (a == c).all() for a, c in zip(ocov, gcov[:ncheck]))
    print(tag, f"windows {nwin} L{L} D{D} e{e} banded {banded}: cons equal {same}/{ncheck} cov equal {csame}/{ncheck} "
          f"status {np.unique(st, return_counts=True)} e2e {dt*1e3:.1f} ms -> {nwin/dt:.0f} windows/s", flush=True)
    # kernel-only timing
    pb.upload(); torch.cuda.synchronize()
    for _ in range(2):
        t = time.time(); pb.launch(); torch.cuda.synchronize(); dt = time.time() - t
    print(tag, f"kernel only {dt*1e3:.1f} ms -> {nwin/dt:.0f} windows/s", flush=True)
    pb.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run("C ", 64, 500, 8, 0.05, False)
    run("A-full", 64, 500, 32, 0.15, False)
    run("A-band", 64, 500, 32, 0.15, True)
    run("Q-full", 64, 400, 20, 0.12, False, q=True)
    run("A-full-2k", 2368, 500, 32, 0.15, False, mem=40 << 30)
    run("A-band-2k", 2368, 500, 32, 0.15, True, mem=40 << 30)
