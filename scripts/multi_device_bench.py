"""Single-process, multi-device mode (racon's own threading model, src/cuda/cudapolisher.cpp:228-240, 336-345):
ONE process drives `batches` batch processors on every visible GPU through b200poa_polisher_create(devices=[...]).
Prints windows/s end to end (host buffers) for 1..N devices; compare with the torchrun numbers of bench.py."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows

ap = argparse.ArgumentParser()
ap.add_argument("--windows-per-gpu", type=int, default=10000)
ap.add_argument("--batches", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--banded", type=int, default=1)
args = ap.parse_args()
ndev = torch.cuda.device_count()
out = []
for n in sorted({1, 2, 4, 8} & set(range(1, ndev + 1)) | {ndev}):
    b = synth_windows(args.windows_per_gpu * n, 500, 32, 0.15, seed=12345)
    pol = api.Polisher(devices=list(range(n)), batches_per_device=args.batches, mem_per_batch=(12 << 30), banded=bool(args.banded))
    chunk = max(256, args.windows_per_gpu // args.batches)
    arena = api.WindowArena.from_batch(b)  # pinned columnar arena: batches upload straight from it
    bufs = None
    for _ in range(2):
        r = pol.polish_arena(arena, max_windows_per_round=chunk, out=bufs)
        bufs = (r[0], r[1], r[2].astype(np.uint8), r[3])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cons, clen, polished, status = pol.polish_arena(arena, max_windows_per_round=chunk, out=bufs)
    dt = time.perf_counter() - t0
    pol.close()
    arena.close()
    out.append({"devices": n, "windows": b.n_windows, "e2e_windows_per_s": b.n_windows * args.steps / dt,
                "unpolished": int((~polished).sum()), "h2d_bytes": pol.last["h2d_bytes"], "d2h_bytes": pol.last["d2h_bytes"]})
    print(json.dumps(out[-1]), flush=True)
