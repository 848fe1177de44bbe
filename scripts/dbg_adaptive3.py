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
pick = [int(x) for x in sys.argv[1:]] or [16, 35]
sub = WindowBatch.from_lists([[(s, None, 0, len(b.window(w)[0][0]) - 1) for s in b.window(w)[0]] for w in pick])
e5 = Emu().polish(sub, identity_order(sub), M, X, G, max_nodes=4092, max_edges=24000, band=512, stride=8192)
g5 = gpu(sub, banded=True, band_width=512)
print("subset", pick, "gpu==emu", [x == y for x, y in zip(g5[0], e5[0])])
