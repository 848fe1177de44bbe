import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import identity_order
from emu_lib import Emu
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows, edit_distance, WindowBatch
M, X, G = 3, -5, -4
def gpu(b, **kw):
    pb = api.PoaBatch(max_gpu_mem=2 << 30, **kw)
    for w in range(b.n_windows):
        seqs = b.window(w)[0]
        assert pb.add_poa_group([(s, None) for s in seqs])[0] == 0
    pb.generate_poa(); out = pb.get_consensus(); pb.close(); return out
b = synth_windows(48, 900, 24, 0.12, seed=7)
ib = identity_order(b)
e5 = Emu().polish(b, ib, M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
for fill in (None, "0", "255", "170", None):
    if fill is None: os.environ.pop("B200POA_SLAB_FILL", None)
    else: os.environ["B200POA_SLAB_FILL"] = fill
    g5 = gpu(b, banded=True, band_width=512)
    bad = [w for w in range(b.n_windows) if g5[0][w] != e5[0][w]]
    print("slab fill", fill, "mismatching", bad)
