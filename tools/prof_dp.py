import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from dp_cases import reference_merges, random_tree
eng = famsa_b200.Engine(0)
rng = np.random.default_rng(0)
n, L = int(sys.argv[1]), int(sys.argv[2])
codes, off, lens = seqio.synth_family(n, L, 5, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o)+int(l)]) for o, l in zip(off, lens)]
g, recs = reference_merges(seqs, random_tree(n, rng, 0.1), threads=(1,))
jobs = [r["job"] for r in recs]
res = eng.dp_align_batch(jobs, g)
print(eng.dp_last_timing())
