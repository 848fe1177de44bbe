timeout 600 python - > gpurun_out/r2_g47_pool.log 2>&1 <<'PY'
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import overlap_fixture, random_pairs
from racon_gpu_b200.aligner import AlignerPool, pack_pairs, pinned
fx = overlap_fixture()
work = {"lambda x8": [(f["q"], f["t"]) for f in fx] * 8, "lambda x64": [(f["q"], f["t"]) for f in fx] * 64,
        "40 kb x480": random_pairs(4242, [(40000, 0.10)] * 24) * 20}
for name, pairs in work.items():
    q, qo, t, to = pack_pairs(pairs)
    with pinned(q, t):
        for nb in (1, 3):
            pool = AlignerPool(devices=(0,), batches_per_device=nb, max_gpu_memory_per_batch=20 << 30)
            best = 1e9
            for it in range(5):
                t0 = time.perf_counter(); ed, buf, off, ln, info = pool.align(q, qo, t, to); dt = time.perf_counter() - t0
                if it >= 2: best = min(best, dt)
            pool.close()
            print(json.dumps({"workload": name, "pairs": len(pairs), "batches": nb, "wall_ms": round(best * 1e3, 2), "overlaps_per_s": round(len(pairs) / best), "kernel_ms_sum": round(info["kernel_ms"], 1), "cells": info["cells"]}))
PY
