import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import identity_order
from emu_lib import Emu
from oracle_lib import Oracle
from test_emu import deletion_windows
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows, edit_distance
M, X, G = 3, -5, -4
MEM = 6 << 30
def gpu(b, order_is_identity=True, **kw):
    pb = api.PoaBatch(max_gpu_mem=MEM, **kw)
    for w in range(b.n_windows):
        seqs = b.window(w)[0]
        assert pb.add_poa_group([(s, None) for s in seqs])[0] == 0
    pb.generate_poa(); out = pb.get_consensus(); pb.close(); return out
e = Emu(); o = Oracle()
# 1. ordinary windows: static vs adaptive (FULLW in the adaptive layout, no retries expected)
a = synth_windows(32, 500, 32, 0.15, seed=83)
ia = identity_order(a)
gs = gpu(a, banded=True); ga = gpu(a, banded="adaptive")
print("1. ordinary: static==adaptive:", sum(x == y for x, y in zip(gs[0], ga[0])), "of", a.n_windows, "status", np.unique(ga[2]))
# 2. static band 512 on 900-bp windows: the multi-chunk banded path vs its emulation
b = synth_windows(16, 900, 24, 0.12, seed=7)
ib = identity_order(b)
for bw in (512, 384):
    g5 = gpu(b, banded=True, band_width=bw)
    e5 = e.polish(b, ib, M, X, G, max_nodes=4092, max_edges=24000, band=bw, stride=8192)
    print(f"2. band {bw}: gpu==emu", sum(x == y for x, y in zip(g5[0], e5[0])), "of", b.n_windows, "status", np.unique(g5[2]), np.unique(e5[2]))
# 3. deletion windows
d = deletion_windows(); idd = identity_order(d)
oc, _, _ = o.polish(d, idd, M, X, G, tgs=False, trim=False, threads=8, stride=8192)
gd = gpu(d, banded="adaptive")
ed = e.polish(d, idd, M, X, G, max_nodes=4092, max_edges=24000, band=-256, stride=8192)
print("3. deletion adaptive: gpu==oracle", [edit_distance(x, y) for x, y in zip(gd[0], oc)], "emu==oracle", sum(x == y for x, y in zip(ed[0], oc)))
g512 = gpu(d, banded=True, band_width=512)
e512 = e.polish(d, idd, M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
print("3b. deletion static 512: gpu==emu", sum(x == y for x, y in zip(g512[0], e512[0])), "status", np.unique(g512[2]))
