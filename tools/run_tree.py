"""Times famsa_prof_align_tree (whole progressive alignment, one call) on the committed hemopexin trees and on a
synthetic family, and a single wide merge through famsa_prof_merge_batch.  usage: run_tree.py [n L]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from famsa_b200.binding import PROF_LEAF
from dp_cases import random_tree

G = os.path.join(ROOT, "tests", "golden")
sm = np.load(os.path.join(G, "adeno_upgma_merges.npz"))["score_matrix"]
eng = famsa_b200.Engine(0)
out = []

def run(name, seqs, merges, gaps, reps=4):
    codes, off, lens = seqio.pack([seqio.encode(s) for s in seqs])
    eng.upload(codes, off, lens); eng.prof_set_scoring(sm)
    best = None
    for _ in range(reps):
        root, _, st = eng.align_tree(merges, gaps, want_paths=False)
        eng.prof_drop([root])
        if best is None or st["wall_ms"] < best["wall_ms"]: best = st
    best["tree"] = name; best["n"] = len(seqs); best["gcells_per_s"] = best["cells"] / best["wall_ms"] / 1e6
    out.append(best); print(json.dumps(best), flush=True)

for f in ("hemopexin_medoid_sl", "hemopexin_sl"):
    z = np.load(os.path.join(G, f + ".npz"))
    run(f, [str(s) for s in z["seqs"]], z["merges"], z["gaps"])
gaps = np.load(os.path.join(G, "hemopexin_sl.npz"))["gaps"]
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 400)
codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
run(f"synthetic {n}x{L} random tree", seqs, np.array(random_tree(n, np.random.default_rng(1), 0.05)), gaps)
run(f"synthetic {n}x{L} chain tree", seqs, np.array(random_tree(n, np.random.default_rng(1), 1.0)), gaps, reps=2)
# single wide merges: a leaf joining the root profile of a 1000-sequence tree, and two halves of the family
codes, off, lens = seqio.synth_family(1200, L, 19, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
eng.upload(codes, off, lens); eng.prof_set_scoring(sm)
def subtree(ids, seed):
    """aligns the sequences `ids` with a random tree level by level, returns the resident root id and its width"""
    from famsa_b200.schedule import ready_levels
    m = random_tree(len(ids), np.random.default_rng(seed), 0.05)
    node = {i: PROF_LEAF | ids[i] for i in range(len(ids))}; width = {i: int(lens[ids[i]]) for i in range(len(ids))}
    for lvl in ready_levels(len(ids), m):
        pairs = [(node.pop(m[k][0]), node.pop(m[k][1])) for k in lvl]
        pid, res = eng.prof_merge_batch(pairs, gaps, [(width[m[k][0]], width[m[k][1]]) for k in lvl])
        for k, p, r in zip(lvl, pid, res):
            node[len(ids) + k] = p; width[len(ids) + k] = len(r["path"])
    r = len(ids) + len(m) - 1
    return node[r], width[r]
for name, build in (("leaf x profile(600 seqs)", lambda: (subtree(list(range(600)), 1), (PROF_LEAF | 1199, int(lens[1199])))),
                    ("profile(600) x profile(599)", lambda: (subtree(list(range(600)), 1), subtree(list(range(600, 1199)), 2)))):
    best = 1e9
    for _ in range(3):
        (a, wa), (b, wb) = build()
        t = time.time(); pid, res = eng.prof_merge_batch([(a, b)], gaps, [(wa, wb)]); dt = time.time() - t
        tot, con = eng.prof_last_timing(); _, ker, cells = eng.dp_last_timing()
        best = min(best, ker); eng.prof_drop(pid)
    rec = dict(merge=name, w=(wa, wb), cells=cells, dp_kernel_ms=best, gcells_per_s=cells / best / 1e6)
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tree_times.json"), "w"), indent=1)
