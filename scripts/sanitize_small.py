"""Tiny run for compute-sanitizer: full-span + partial-span windows, both band modes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows
from common import partial_span_windows

for b in (partial_span_windows(n=6, length=300, depth=10), synth_windows(6, 300, 10, 0.1, seed=3)):
    for banded in (False, True):
        pb = api.PoaBatch(max_gpu_mem=1 << 30, banded=banded)
        pb.add_windows(b)
        pb.generate_poa()
        cons, cov, st = pb.get_consensus()
        pb.close()
        print(banded, [len(c) for c in cons], st.tolist())
