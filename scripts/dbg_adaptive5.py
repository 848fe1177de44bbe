import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import identity_order
from emu_lib import Emu
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows, edit_distance, WindowBatch
M, X, G = 3, -5, -4
def gpu(wins, **kw):
    pb = api.PoaBatch(max_gpu_mem=2 << 30, **kw)
    for seqs in wins:
        assert pb.add_poa_group([(s, None) for s in seqs])[0] == 0
    pb.generate_poa(); out = pb.get_consensus(); pb.close(); return out
b = synth_windows(48, 900, 24, 0.12, seed=7)
wins = [b.window(w)[0] for w in range(48)]
e5 = Emu().polish(b, identity_order(b), M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)[0]
def check(idx, label, **kw):
    g = gpu([wins[i] for i in idx], banded=True, band_width=512, **kw)[0]
    print(label, "bad (original ids):", [idx[k] for k in range(len(idx)) if g[k] != e5[idx[k]]])
check(list(range(48)), "all 48")
check(list(range(47, -1, -1)), "reversed")
check([16, 35, 11], "16,35 + the longest-read window 11")
check([16, 11], "16 + 11")
check([16], "16 alone")
check([16, 35], "16,35")
# which read of window 16 first diverges when window 11 sets the geometry?
for d in range(3, 26):
    g = gpu([wins[16][:d], wins[11]], banded=True, band_width=512)
    sub = WindowBatch.from_lists([[(s, None, 0, len(wins[16][0]) - 1) for s in wins[16][:d]]])
    ee = Emu().polish(sub, identity_order(sub), M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
    if g[0][0] != ee[0][0] or not (g[1][0] == ee[1][0]).all():
        print("window 16 diverges with", d, "sequences; last read length", len(wins[16][d - 1]))
        break
