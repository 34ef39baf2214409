import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
codes, offsets, lens = seqio.synth_family(n, 400, seed=1)
eng = famsa_b200.Engine(0)
eng.upload(codes, offsets, lens)
for r in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t = time.time(); ef, et, ed, order = eng.prim(0); wall = time.time() - t
    tot, main, p = eng.last_timing()
    print(f"n={n}: famsa_lcs_prim wall {wall*1e3:.1f} ms (device total {tot:.1f} ms, of which LCS tile kernels {main:.1f} ms); edges {len(ef)}, sum dist {ed.sum():.6f}")
