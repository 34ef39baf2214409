"""Whole progressive alignment on several GPUs with resident profiles, sharded by subtree (famsa_b200.schedule).
torchrun --nproc-per-node N tools/run_tree_resident_mgpu.py NSEQ LEN    (object collectives over gloo)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch.distributed as dist
import famsa_b200
from famsa_b200 import schedule, seqio
from oracle import pyoracle
from dp_cases import assemble_rows, random_tree, resident_progressive_alignment

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
n, L = int(sys.argv[1]), int(sys.argv[2])
codes, off, lens = seqio.synth_family(n, L, 17, sort_desc=False)
seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
merges = random_tree(n, np.random.default_rng(1), 0.05)
dp = pyoracle.RefDp(n); sm = dp.score_matrix(); g = dp.gaps(); dp.close()
eng = famsa_b200.Engine(local)
eng.upload(codes, off, lens)
eng.prof_set_scoring(sm)
for rep in range(2):
    if world > 1:
        dist.barrier()
    t = time.time()
    results, root = schedule.sharded_resident_alignment(eng, dist if world > 1 else None, rank, world, n, lens, merges, g)
    wall = time.time() - t
    if root is not None:
        eng.prof_drop([root])
owner, frontier = schedule.subtree_shards(n, merges, world)
parts = [None] * world
if world > 1:
    dist.all_gather_object(parts, (wall, {k: (r["path"], r["swapped"], r["total"]) for k, r in results.items()}))
else:
    parts = [(wall, {k: (r["path"], r["swapped"], r["total"]) for k, r in results.items()})]
if rank == 0:
    merged = {}
    for _, p in parts:
        merged.update(p)
    res = {k: dict(path=v[0], swapped=v[1], total=v[2]) for k, v in merged.items()}
    rows = assemble_rows(seqs, merges, res)
    # single-GPU resident run of the same tree as the cross-check
    rows1, res1, root1 = resident_progressive_alignment(eng, seqs, merges, g, sm)
    eng.prof_drop([root1])
    print(json.dumps({"config": f"{n} x {L} aa, {len(merges)} merges, {world} GPUs, subtree sharding",
                      "wall_s_per_rank": [w for w, _ in parts], "merges_per_rank": [sum(1 for o in owner if o == r) for r in range(world)],
                      "top_merges_on_rank0": sum(1 for o in owner if o == -1), "subtree_roots_moved": sum(1 for v, r in frontier if r != 0 and v >= n),
                      "identical_to_single_gpu": bool(rows == rows1 and all(res[k]["total"] == res1[k]["total"] for k in range(len(merges)))),
                      "final_width": len(rows[0])}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
