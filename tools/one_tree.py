"""One famsa_prof_align_tree run of a committed hemopexin tree (for ncu launch lists).  usage: one_tree.py hemopexin_sl"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import famsa_b200
from famsa_b200 import seqio
G = os.path.join(ROOT, "tests", "golden")
sm = np.load(os.path.join(G, "adeno_upgma_merges.npz"))["score_matrix"]
z = np.load(os.path.join(G, (sys.argv[1] if len(sys.argv) > 1 else "hemopexin_sl") + ".npz"))
eng = famsa_b200.Engine(0)
codes, off, lens = seqio.pack([seqio.encode(str(s)) for s in z["seqs"]])
eng.upload(codes, off, lens); eng.prof_set_scoring(sm)
root, _, st = eng.align_tree(z["merges"], z["gaps"], want_paths=False); eng.prof_drop([root])
torch.cuda.cudart().cudaProfilerStart()
root, _, st = eng.align_tree(z["merges"], z["gaps"], want_paths=False)
torch.cuda.cudart().cudaProfilerStop()
print(st)
