"""Level-synchronous progressive alignment of a synthetic family with the DP on the GPU and the reference's own
ConstructProfile on the host (evidence run: parity at scale + per-level timing).  usage: run_tree_dp.py N L"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from famsa_b200.schedule import ready_levels
from oracle import pyoracle
from dp_cases import random_tree, reference_merges, driven_progressive_alignment

n, L = int(sys.argv[1]), int(sys.argv[2])
codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
merges = random_tree(n, np.random.default_rng(1), 0.05)
eng = famsa_b200.Engine(0)
stats = dict(gpu_s=0.0, kernel_ms=0.0, cells=0, levels=0, max_level=0)
def level(jobs, g):
    t = time.time(); r = eng.dp_align_batch(jobs, g, want_dirs=True); stats["gpu_s"] += time.time() - t
    tot, k, c = eng.dp_last_timing(); stats["kernel_ms"] += k; stats["cells"] += c; stats["levels"] += 1
    stats["max_level"] = max(stats["max_level"], len(jobs))
    return r
t = time.time(); rows, total = driven_progressive_alignment(seqs, merges, level); t_gpu_path = time.time() - t
t = time.time(); g, recs = reference_merges(seqs, merges, threads=(1,)); t_ref = time.time() - t
ok = rows == recs[-1]["rows"] and total == recs[-1]["total"]
print(json.dumps({"config": f"{n} x {L} aa synthetic family, random guide tree, {len(merges)} merges in {stats['levels']} levels",
                  "alignment_identical_to_reference": bool(ok), "final_width": len(rows[0]), "dp_cells": stats["cells"],
                  "gpu_dp_kernel_ms": stats["kernel_ms"], "gpu_dp_calls_s": stats["gpu_s"], "largest_level": stats["max_level"],
                  "gpu_path_total_s(incl. reference ConstructProfile + python)": t_gpu_path, "reference_cpu_1thread_s": t_ref,
                  "gpu_dp_cells_per_s_kernel": stats["cells"] / (stats["kernel_ms"] / 1e3)}))
