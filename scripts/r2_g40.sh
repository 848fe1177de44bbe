set -x
nvidia-smi -L > gpurun_out/r2_g40_gpus.log 2>&1
timeout 600 python - > gpurun_out/r2_g40_pool2.log 2>&1 <<'PY'
import sys, time, json, hashlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import overlap_fixture
from racon_gpu_b200.aligner import AlignerPool, pack_pairs, pinned
fx = overlap_fixture()
rep = 64
q, qo, t, to = pack_pairs([(f["q"], f["t"]) for f in fx] * rep)
with pinned(q, t):
    for devices, nb in (((0,), 3), ((0, 1), 2), ((0, 1), 3)):
        pool = AlignerPool(devices=devices, batches_per_device=nb, max_gpu_memory_per_batch=20 << 30)
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter(); ed, buf, off, ln, info = pool.align(q, qo, t, to); dt = time.perf_counter() - t0
            if it: best = min(best, dt)
        bad = 0
        for k in range(len(fx) * rep):
            f = fx[k % len(fx)]
            if ed[k] != f["score"] or hashlib.sha256(buf[off[k]:off[k] + ln[k]].tobytes()).hexdigest() != f["cigar_sha"]: bad += 1
        pool.close()
        print(json.dumps({"devices": devices, "batches_per_device": nb, "pairs": len(qo) - 1, "wall_ms": best * 1e3, "overlaps_per_s": (len(qo) - 1) / best, "wrong": bad}))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_g40_bench_n2.json 2> gpurun_out/r2_g40_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_device" 2>&1 | tail -3 > gpurun_out/r2_g40_multidev_test.log
