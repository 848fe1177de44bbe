import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import identity_order
from emu_lib import Emu
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows
M, X, G = 3, -5, -4
def gpu(wins, **kw):
    pb = api.PoaBatch(max_gpu_mem=2 << 30, **kw)
    for seqs in wins:
        assert pb.add_poa_group([(s, None) for s in seqs])[0] == 0
    pb.generate_poa(); out = pb.get_consensus(); pb.close(); return out
b = synth_windows(48, 900, 24, 0.12, seed=7)
wins = [b.window(w)[0] for w in range(48)]
e5 = Emu().polish(b, identity_order(b), M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)[0]
tot = 0
for rep in range(12):
    g = gpu(wins, banded=True, band_width=512)[0]
    bad = [w for w in range(48) if g[w] != e5[w]]
    tot += len(bad)
    print("rep", rep, "bad", bad)
print(os.environ.get("B200POA_LIB", "in-tree"), "TOTAL mismatches", tot)
