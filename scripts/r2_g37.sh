set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g37_aln_tests.log
L=gpurun_out/r2_g37_aln_bench.log; : > $L
for view in 0 1 0 1; do
  for rep in 8 64; do echo "== view $view rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 --view $view >> $L 2>&1; done
done
timeout 600 python - > gpurun_out/r2_g37_pool.log 2>&1 <<'PY'
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import overlap_fixture
from racon_gpu_b200.aligner import AlignerPool, pack_pairs, pinned
fx = overlap_fixture()
for rep in (8, 64):
    q, qo, t, to = pack_pairs([(f["q"], f["t"]) for f in fx] * rep)
    with pinned(q, t):
        for nb in (1, 2, 3):
            pool = AlignerPool(devices=(0,), batches_per_device=nb, max_gpu_memory_per_batch=20 << 30)
            best = 1e9
            for it in range(4):
                t0 = time.perf_counter(); ed, buf, off, ln, info = pool.align(q, qo, t, to); dt = time.perf_counter() - t0
                if it: best = min(best, dt)
            pool.close()
            print(json.dumps({"rep": rep, "batches": nb, "wall_ms": best * 1e3, "overlaps_per_s": (len(qo) - 1) / best, "kernel_ms_sum": info["kernel_ms"]}))
PY
