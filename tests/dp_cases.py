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
                     gaps=None, want_merged: bool = False):
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
        if want_merged:                          # the tables ConstructProfile built (profile.cpp:784-1002)
            recs[-1]["merged"] = dp.tables(m)
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


def resident_progressive_alignment(engine, seqs, merges, gaps, score_matrix, on_level=None):
    """Level-synchronous progressive alignment with every profile resident on the GPU (famsa_prof_merge_batch):
    leaves come from the uploaded sequences, merged tables never leave HBM, the host receives one path per merge
    and applies its gap runs to the member rows (what FinalizeGaps does, profile.cpp:1053-1104).
    on_level(level_merge_indices, merged_ids, results) is called after each level, before the next consumes them.
    Returns ({seq_no: gapped string}, results per merge, root id)."""
    from famsa_b200 import seqio
    from famsa_b200.binding import PROF_LEAF
    from famsa_b200.schedule import ready_levels
    n = len(seqs)
    codes, off, lens = seqio.pack([seqio.encode(s) for s in seqs])
    engine.upload(codes, off, lens)
    engine.prof_set_scoring(score_matrix)
    node = {i: PROF_LEAF | i for i in range(n)}
    width = {i: len(seqs[i]) for i in range(n)}
    rows = {i: {i: np.frombuffer(seqs[i].encode(), dtype=np.uint8)} for i in range(n)}
    results = [None] * len(merges)
    for lvl in ready_levels(n, merges):
        pairs = [(node.pop(merges[k][0]), node.pop(merges[k][1])) for k in lvl]
        ids, res = engine.prof_merge_batch(pairs, gaps, [(width[merges[k][0]], width[merges[k][1]]) for k in lvl])
        for k, pid, r in zip(lvl, ids, res):
            a, b = merges[k]
            node[n + k] = pid
            width[n + k] = len(r["path"])
            ra, rb = rows.pop(a), rows.pop(b)
            rrows, crows = (rb, ra) if r["swapped"] else (ra, rb)
            out = {}
            for members, gapdir in ((rrows, 1), (crows, 2)):
                keep = r["path"] != gapdir
                for no, row in members.items():
                    g = np.full(len(r["path"]), ord("-"), dtype=np.uint8)
                    g[keep] = row
                    out[no] = g
            rows[n + k] = out
            results[k] = r
        if on_level:
            on_level(lvl, ids, res)
    root = n + len(merges) - 1
    return {no: row.tobytes().decode() for no, row in rows[root].items()}, results, node[root]


def assemble_rows(seqs, merges, results):
    """Final alignment {seq_no: gapped string} from the per-merge paths (what FinalizeGaps does on the host):
    H steps are gap columns in the members of the DP's row profile, V steps in those of the column profile."""
    n = len(seqs)
    rows = {i: {i: np.frombuffer(seqs[i].encode(), dtype=np.uint8)} for i in range(n)}
    for k, (a, b) in enumerate(merges):
        r = results[k]
        ra, rb = rows.pop(a), rows.pop(b)
        rrows, crows = (rb, ra) if r["swapped"] else (ra, rb)
        out = {}
        for members, gapdir in ((rrows, 1), (crows, 2)):
            keep = r["path"] != gapdir
            for no, row in members.items():
                g = np.full(len(r["path"]), ord("-"), dtype=np.uint8)
                g[keep] = row
                out[no] = g
        rows[n + k] = out
    return {no: row.tobytes().decode() for no, row in rows[n + len(merges) - 1].items()}


class OracleEngine:
    """CPU stand-in for Engine's resident-profile calls, built on the oracle (tests of the multi-rank host logic run
    without a GPU): same ids / leaf handles / consume-on-merge semantics as famsa_prof_*."""
    def __init__(self):
        self.tab = {}
        self.next = 0

    def upload(self, codes, off, lens):
        self.seqs = [np.asarray(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]

    def prof_set_scoring(self, sm):
        self.sm = np.asarray(sm, dtype=np.int64)

    def _tables(self, h, gaps):
        from famsa_b200 import profiles
        from famsa_b200.binding import PROF_LEAF
        if h & PROF_LEAF:
            return profiles.tables_from_rows(self.seqs[h & ~PROF_LEAF][None, :], self.sm, gaps)
        return self.tab.pop(h)

    def prof_merge_batch(self, pairs, gaps, widths):
        ids, out = [], []
        for a, b in pairs:
            ta, tb = self._tables(a, gaps), self._tables(b, gaps)
            r = pyoracle.dp_align(*ta, *tb, gaps)
            rp, cp = (tb, ta) if r["swapped"] else (ta, tb)
            s, c, _, _ = pyoracle.dp_construct(rp, cp, r["path"], gaps)
            self.tab[self.next] = (s, c, ta[2] + tb[2])
            ids.append(self.next); self.next += 1
            out.append(dict(path=r["path"], total=r["total"], last=r["last"], swapped=r["swapped"], variant=r["variant"]))
        return ids, out

    def prof_get(self, pid, tables=True):
        s, c, k = self.tab[pid]
        return (s, c, k) if tables else (s.shape[0] - 1, k)

    def prof_put(self, profs):
        ids = []
        for p in profs:
            self.tab[self.next] = p
            ids.append(self.next); self.next += 1
        return ids

    def prof_drop(self, ids):
        for i in ids:
            del self.tab[int(i)]
