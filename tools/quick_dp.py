"""Quick timing of the DP batch path (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from dp_cases import reference_merges, random_tree

eng = famsa_b200.Engine(0)
def bench(name, seqs, merges, rescale=None):
    t = time.time(); g, recs = reference_merges(seqs, merges, n_seqs_for_rescale=rescale, threads=(1,)); tref = time.time() - t
    jobs = [r["job"] for r in recs]
    cells = sum((j[0].shape[0]-1)*(j[3].shape[0]-1) for j in jobs)
    for rep in range(2):
        t = time.time(); res = eng.dp_align_batch(jobs, g); wall = time.time() - t
        tot, kern, c = eng.dp_last_timing()
    ok = all(r["total"] == rec["total"] for r, rec in zip(res, recs))
    print(f"{name}: {len(jobs)} merges {cells/1e6:.1f} Mcells | kernel {kern:.2f} ms = {cells/kern/1e3:.0f} Mc/s | e2e {wall*1e3:.1f} ms = {cells/wall/1e6:.0f} Mc/s | ok={ok}")

z = np.load(os.path.join(ROOT, "tests/golden/hemopexin_medoid_sl.npz"))
bench("hemopexin C4 (one batch)", [str(s) for s in z["seqs"]], [tuple(int(x) for x in m) for m in z["merges"]])
rng = np.random.default_rng(0)
for n, L in ((400, 400), (64, 1500)):
    codes, off, lens = seqio.synth_family(n, L, 5, sort_desc=False)
    seqs = [seqio.decode(codes[int(o):int(o)+int(l)]) for o, l in zip(off, lens)]
    bench(f"synthetic {n}x{L}", seqs, random_tree(n, rng, 0.1))
