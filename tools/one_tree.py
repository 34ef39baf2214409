"""One famsa_prof_align_tree run of a committed hemopexin tree (for ncu launch lists).  usage: one_tree.py hemopexin_sl"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import famsa_b200
from famsa_b200 import seqio
G = os.path.join(ROOT, "tests", "golden")
sm = np.load(os.path.join(G, "adeno_upgma_merges.npz"))["score_matrix"]
name = sys.argv[1] if len(sys.argv) > 1 else "hemopexin_sl"
eng = famsa_b200.Engine(0)
if name.startswith("synthetic"):                       # synthetic:N:L[:caterpillar]
    from dp_cases import random_tree
    parts = name.split(":")
    n, L = int(parts[1]), int(parts[2])
    codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
    z = {"merges": np.array(random_tree(n, np.random.default_rng(1), float(parts[3]) if len(parts) > 3 else 0.05)),
         "gaps": np.load(os.path.join(G, "hemopexin_sl.npz"))["gaps"]}
else:
    z = np.load(os.path.join(G, name + ".npz"))
    codes, off, lens = seqio.pack([seqio.encode(str(s)) for s in z["seqs"]])
eng.upload(codes, off, lens); eng.prof_set_scoring(sm)
root, _, st = eng.align_tree(z["merges"], z["gaps"], want_paths=False); eng.prof_drop([root])
import ctypes as C
if os.environ.get('FAMSA_FUSED_TIMING'):
    eng.lib.famsa_debug_fused_phases((C.c_double * 8)())
torch.cuda.cudart().cudaProfilerStart()
root, _, st = eng.align_tree(z["merges"], z["gaps"], want_paths=False)
torch.cuda.cudart().cudaProfilerStop()
print(st)
import ctypes as C
if os.environ.get('FAMSA_FUSED_TIMING'):
    out = (C.c_double * 8)()
    eng.lib.famsa_debug_fused_phases(out)
    names = ['leaf', 'prep', 'fill', 'dirs->smem', 'trace', 'construct', 'barrier_wait', 'idle_before_launch']
    print({k: round(v / 1e3, 1) for k, v in zip(names, out)}, 'us total; levels:', st['n_batches'])
