"""End-to-end probe: windows/s through api.Polisher.polish for several (batch processors, rounds) settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows
nwin = int(os.environ.get("NWIN", "10000"))
b = synth_windows(nwin, 500, 32, 0.15, seed=12345)
for batches, rounds in ((1, 1), (2, 1), (4, 1), (2, 2), (4, 2), (3, 1), (8, 1)):
    pol = api.Polisher(devices=[0], batches_per_device=batches, mem_per_batch=(48 << 30) // batches, banded=True)
    chunk = max(256, int(-(-nwin // (batches * rounds))))
    out = None
    for _ in range(2):
        r = pol.polish(b, max_windows_per_round=chunk, out=out)
        out = (r[0], r[1], r[2].astype(np.uint8), r[3])
    t0 = time.perf_counter()
    K = 4
    for _ in range(K):
        pol.polish(b, max_windows_per_round=chunk, out=out)
    dt = (time.perf_counter() - t0) / K
    print(f"batches {batches} rounds {rounds} chunk {chunk}: {1e3*dt:.1f} ms/step  {nwin/dt:.0f} windows/s  launches {pol.last['kernel_launches']}", flush=True)
    pol.close()
