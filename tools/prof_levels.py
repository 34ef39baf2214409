"""Per-level breakdown of a resident progressive alignment.  usage: prof_levels.py N L"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from oracle import pyoracle
from dp_cases import random_tree, resident_progressive_alignment

n, L = int(sys.argv[1]), int(sys.argv[2])
codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
merges = random_tree(n, np.random.default_rng(1), 0.05)
dp = pyoracle.RefDp(n); sm = dp.score_matrix(); g = dp.gaps(); dp.close()
eng = famsa_b200.Engine(0)
orig = eng.prof_merge_batch
rows = []
def timed(pairs, gaps, widths):
    t = time.time(); r = orig(pairs, gaps, widths); call = time.time() - t
    a, b = eng.prof_last_timing(); _, k, c = eng.dp_last_timing()
    rows.append((len(pairs), c, max(max(w) for w in widths), max(min(w) for w in widths), k, a, call * 1e3))
    return r
eng.prof_merge_batch = timed
for rep in range(2):
    rows.clear()
    _, _, root = resident_progressive_alignment(eng, seqs, merges, g, sm)
    eng.prof_drop([root])
print("level merges cells maxW max(minW) dp_kernel_ms device_ms call_ms Gcells/s")
for i, r in enumerate(rows):
    print(i, r[0], r[1], r[2], r[3], f"{r[4]:.3f} {r[5]:.3f} {r[6]:.3f} {r[1] / r[4] / 1e6:.2f}")
