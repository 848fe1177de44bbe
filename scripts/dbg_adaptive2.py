import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import identity_order
from emu_lib import Emu
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows, edit_distance
M, X, G = 3, -5, -4
MEM = 6 << 30
def gpu(b, **kw):
    pb = api.PoaBatch(max_gpu_mem=MEM, **kw)
    for w in range(b.n_windows):
        seqs = b.window(w)[0]
        assert pb.add_poa_group([(s, None) for s in seqs])[0] == 0
    pb.generate_poa(); out = pb.get_consensus(); pb.close(); return out
e = Emu()
b = synth_windows(48, 900, 24, 0.12, seed=7)
ib = identity_order(b)
e5 = e.polish(b, ib, M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
for rb in (None, "70000", "2300"):
    if rb: os.environ["B200POA_RING_BYTES"] = rb
    g5 = gpu(b, banded=True, band_width=512)
    bad = [w for w in range(b.n_windows) if g5[0][w] != e5[0][w]]
    print("band 512 ring_bytes", rb, "mismatching windows", bad, [edit_distance(g5[0][w], e5[0][w]) for w in bad])
os.environ.pop("B200POA_RING_BYTES", None)
# depth sweep on the first mismatching window: at which read does it diverge?
g5 = gpu(b, banded=True, band_width=512)
bad = [w for w in range(b.n_windows) if g5[0][w] != e5[0][w]]
if bad:
    w = bad[0]
    seqs = b.window(w)[0]
    from racon_gpu_b200.windows import WindowBatch
    for d in range(3, len(seqs) + 1):
        sub = WindowBatch.from_lists([[(s, None, 0, len(seqs[0]) - 1) for s in seqs[:d]]])
        gg = gpu(sub, banded=True, band_width=512)
        ee = e.polish(sub, identity_order(sub), M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
        same = gg[0] == ee[0] and all((x == y).all() for x, y in zip(gg[1], ee[1]))
        if not same:
            print("window", w, "first divergence with", d, "sequences; read length", len(seqs[d - 1]), "lens", [len(s) for s in seqs[:d]])
            break
