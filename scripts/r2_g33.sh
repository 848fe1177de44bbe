set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g33_aln_tests.log
timeout 600 python - > gpurun_out/r2_g33_bp.log 2>&1 <<'PY'
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from common import overlap_fixture
from racon_gpu_b200.aligner import CUDABatchAligner, pack_pairs
fx = overlap_fixture()
rep = 64
q, qo, t, to = pack_pairs([(f["q"], f["t"]) for f in fx] * rep)
qf = np.asarray([f["q_first"] for f in fx] * rep, dtype=np.int32); tb = np.asarray([f["t_begin"] for f in fx] * rep, dtype=np.int32)
for mode in ("cigars", "cigars+bp", "bp only"):
    al = CUDABatchAligner(device_id=0, max_gpu_memory=32 << 30)
    if mode != "cigars": al.set_window_length(500, skip_cigars=(mode == "bp only"))
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        al.add_overlaps(q, qo, t, to, 0, qf, tb); al.align_all()
        if mode == "bp only":
            st = al.lib.b200aln_batch_sync(al.h)
        else:
            al.cigars()
        info = al.info(); al.reset()
        dt = time.perf_counter() - t0
        if it: best = min(best, dt)
    al.close()
    print(json.dumps({"mode": mode, "wall_ms": best * 1e3, "kernel_ms": info["kernel_ms"], "d2h": info["d2h_bytes"], "overlaps_per_s": len(qf) / best}))
PY
