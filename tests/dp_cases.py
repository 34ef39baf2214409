"""Shared helpers for the HP-2 tests: drive the reference (oracle/_ref) through a guide tree, collecting
for every merge the inputs of CProfile::Align and the reference's outcome (path, total score)."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle


def random_tree(n: int, rng, caterpillar: float = 0.3) -> list[tuple[int, int]]:
    """Random binary merge order over n leaves (ids as in tree_structure)."""
    alive = list(range(n))
    merges = []
    while len(alive) > 1:
        if rng.random() < caterpillar and len(merges):
            a = alive.pop()                      # extend the newest node (deep, chain-like)
            b = alive.pop(int(rng.integers(len(alive))))
        else:
            a = alive.pop(int(rng.integers(len(alive))))
            b = alive.pop(int(rng.integers(len(alive))))
        merges.append((a, b))
        alive.append(n + len(merges) - 1)
    return merges


def reference_merges(seqs: list[str], merges, n_seqs_for_rescale: int | None = None, threads=(1, 2), rng=None,
                     gaps=None):
    """Runs the reference's progressive alignment.  Returns (gaps, records) where each record is a dict with
    the Align inputs (s1,c1,k1,s2,c2,k2), members, and the reference's result (total, rows of the merged
    profile -> path via pyoracle.path_from_rows once the orientation is known)."""
    rng = rng or np.random.default_rng(0)
    n = len(seqs)
    dp = pyoracle.RefDp(n if n_seqs_for_rescale is None else n_seqs_for_rescale)
    if gaps is not None:
        dp.set_gaps(gaps)
    g = dp.gaps()
    nodes = {i: (dp.leaf(seqs[i], i), {i}) for i in range(n)}
    recs = []
    for k, (a, b) in enumerate(merges):
        pa, ma = nodes.pop(a)
        pb, mb = nodes.pop(b)
        s1, c1, k1 = dp.tables(pa)
        s2, c2, k2 = dp.tables(pb)
        m, total = dp.align(pa, pb, int(rng.choice(threads)))
        recs.append(dict(job=(s1, c1, k1, s2, c2, k2), m1=ma, m2=mb, total=total, rows=dp.rows(m)))
        nodes[n + k] = (m, ma | mb)
    for p, _ in nodes.values():
        dp.free(p)
    dp.close()
    return g, recs


def check_against_reference(results, recs):
    """results: list of dicts with path/total/swapped (oracle or GPU), in the order of recs."""
    for k, (r, rec) in enumerate(zip(results, recs)):
        want = pyoracle.path_from_rows(rec["rows"], rec["m1"], rec["m2"], r["swapped"])
        assert r["total"] == rec["total"], f"merge {k}: total {r['total']} != {rec['total']}"
        assert np.array_equal(r["path"], want[:len(r["path"])]) and len(want) == len(r["path"]), f"merge {k}: path differs"


def driven_progressive_alignment(seqs, merges, align_level, n_seqs_for_rescale=None):
    """Level-synchronous progressive alignment in which the DP (direction matrices + corner scores) comes from
    `align_level(jobs, gaps) -> [dict(dirs, last, swapped), ...]` and everything else -- leaf profiles, the merged
    profile construction -- is the reference's own host code (ConstructProfile, profile.cpp:694-1002, through
    oracle/ref_harness.cpp).  Returns the final alignment rows {seq_no: gapped string} and the root total score."""
    from famsa_b200.schedule import ready_levels
    n = len(seqs)
    dp = pyoracle.RefDp(n if n_seqs_for_rescale is None else n_seqs_for_rescale)
    g = dp.gaps()
    nodes = {i: dp.leaf(seqs[i], i) for i in range(n)}
    for lvl in ready_levels(n, merges):
        jobs = []
        for k in lvl:
            a, b = merges[k]
            s1, c1, k1 = dp.tables(nodes[a])
            s2, c2, k2 = dp.tables(nodes[b])
            jobs.append((s1, c1, k1, s2, c2, k2))
        res = align_level(jobs, g)
        for k, r in zip(lvl, res):
            a, b = merges[k]
            nodes[n + k] = dp.construct(nodes.pop(a), nodes.pop(b), r["dirs"], r["last"], r["swapped"])
    root = nodes[n + len(merges) - 1]
    rows = dp.rows(root)
    total = int(dp.lib.ref_profile_total_score(root))
    dp.free(root)
    dp.close()
    return rows, total
