"""Whole progressive alignment of a synthetic family with every profile resident in HBM (famsa_prof_merge_batch):
per level one call, per merge only the path returns.  Compares the assembled alignment with the reference's and
reports per-phase times.  usage: run_tree_resident.py N L [check]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import famsa_b200
from famsa_b200 import seqio
from oracle import pyoracle
from dp_cases import random_tree, reference_merges, resident_progressive_alignment

n, L = int(sys.argv[1]), int(sys.argv[2])
check = len(sys.argv) > 3
codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
merges = random_tree(n, np.random.default_rng(1), 0.05)
dp = pyoracle.RefDp(n); sm = dp.score_matrix(); g = dp.gaps(); dp.close()
eng = famsa_b200.Engine(0)
st = dict(total_ms=0.0, construct_ms=0.0, dp_kernel_ms=0.0, cells=0, levels=0, call_s=0.0, peak_bytes=0)
orig = eng.prof_merge_batch
def timed(pairs, gaps, widths):
    t = time.time(); r = orig(pairs, gaps, widths); st["call_s"] += time.time() - t
    a, b = eng.prof_last_timing(); st["total_ms"] += a; st["construct_ms"] += b
    _, k, c = eng.dp_last_timing(); st["dp_kernel_ms"] += k; st["cells"] += c; st["levels"] += 1
    st["peak_bytes"] = max(st["peak_bytes"], eng.prof_stats()[1])
    return r
eng.prof_merge_batch = timed
for rep in range(2):                     # second pass: warm allocator / buffers
    for k in st: st[k] = 0
    t = time.time(); rows, res, root = resident_progressive_alignment(eng, seqs, merges, g, sm); wall = time.time() - t
    eng.prof_drop([root])
out = {"config": f"{n} x {L} aa synthetic family, random guide tree, {len(merges)} merges in {st['levels']} levels",
       "final_width": len(rows[0]), "dp_cells": st["cells"], "device_ms_all_levels": st["total_ms"],
       "dp_kernel_ms": st["dp_kernel_ms"], "construct_kernel_ms": st["construct_ms"], "merge_batch_calls_s": st["call_s"],
       "python_wall_s(incl. row assembly)": wall, "peak_resident_bytes": st["peak_bytes"],
       "cells_per_s_through_calls": st["cells"] / st["call_s"]}
if check:
    t = time.time(); _, recs = reference_merges(seqs, merges, threads=(1,)); out["reference_cpu_1thread_s"] = time.time() - t
    out["alignment_identical_to_reference"] = bool(rows == recs[-1]["rows"] and res[-1]["total"] == recs[-1]["total"])
print(json.dumps(out))
