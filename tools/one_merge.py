"""One wide merge alone on the GPU (for ncu): a leaf joining the root profile of a 600-sequence tree (SeqProf) and two
half-family profiles (ProfProf).  usage: one_merge.py [seq|prof]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from famsa_b200.binding import PROF_LEAF
from famsa_b200.schedule import ready_levels
from dp_cases import random_tree
G = os.path.join(ROOT, "tests", "golden")
sm = np.load(os.path.join(G, "adeno_upgma_merges.npz"))["score_matrix"]
gaps = np.load(os.path.join(G, "hemopexin_sl.npz"))["gaps"]
which = sys.argv[1] if len(sys.argv) > 1 else "seq"
eng = famsa_b200.Engine(0)
codes, off, lens = seqio.synth_family(1200, 400, 19, sort_desc=False)
eng.upload(codes, off, lens); eng.prof_set_scoring(sm)
def subtree(ids, seed):
    m = random_tree(len(ids), np.random.default_rng(seed), 0.05)
    node = {i: PROF_LEAF | ids[i] for i in range(len(ids))}; width = {i: int(lens[ids[i]]) for i in range(len(ids))}
    for lvl in ready_levels(len(ids), m):
        pairs = [(node.pop(m[k][0]), node.pop(m[k][1])) for k in lvl]
        pid, res = eng.prof_merge_batch(pairs, gaps, [(width[m[k][0]], width[m[k][1]]) for k in lvl])
        for k, p, r in zip(lvl, pid, res):
            node[len(ids) + k] = p; width[len(ids) + k] = len(r["path"])
    r = len(ids) + len(m) - 1
    return node[r], width[r]
a, wa = subtree(list(range(600)), 1)
b, wb = (PROF_LEAF | 1199, int(lens[1199])) if which == "seq" else subtree(list(range(600, 1199)), 2)
import torch
torch.cuda.cudart().cudaProfilerStart()
pid, res = eng.prof_merge_batch([(a, b)], gaps, [(wa, wb)])
torch.cuda.cudart().cudaProfilerStop()
print(which, wa, wb, eng.dp_last_timing())
