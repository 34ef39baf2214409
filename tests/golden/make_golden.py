"""Generates the committed golden fixtures from the reference's own test data.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Outputs (small, committed):
  adeno_fiber_lcs.npz   sequences of test/adeno_fiber/adeno_fiber + the exact LCS lengths pinned by
                        test/adeno_fiber/pid_sq.csv (pid = lcs / min(len), 6 decimals; row = seq0)
                        + the distances of dist_sq.csv (indel075_div_lcs, float, 6 decimals)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from famsa_b200 import seqio  # noqa: E402

REF = "/root/reference/test"


def read_square(path):
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split(",")[1:]
        rows = []
        for line in fh:
            parts = line.rstrip("\n").split(",")
            rows.append([float(x) for x in parts[1:]])
    return header, np.array(rows, dtype=np.float64)


def main():
    ids, seqs = seqio.read_fasta(os.path.join(REF, "adeno_fiber", "adeno_fiber"))
    codes = [seqio.encode(s) for s in seqs]
    lens = np.array([len(c) for c in codes])
    hdr, pid = read_square(os.path.join(REF, "adeno_fiber", "pid_sq.csv"))
    assert hdr == ids and pid.shape == (len(ids), len(ids))
    minlen = np.minimum(lens[:, None], lens[None, :])
    lcs = np.rint(pid * minlen).astype(np.uint16)
    # the 6-decimal print must round-trip, otherwise the fixture would not pin the integer
    assert np.all(np.abs(lcs / minlen - pid) < 6e-7)
    _, dist = read_square(os.path.join(REF, "adeno_fiber", "dist_sq.csv"))
    np.savez_compressed(os.path.join(HERE, "adeno_fiber_lcs.npz"),
                        seqs=np.array(seqs), ids=np.array(ids), lcs=lcs, dist=dist.astype(np.float64))
    print("adeno_fiber_lcs.npz:", lcs.shape, "lcs range", lcs.min(), lcs.max())


def make_dp():
    """HP-2 fixtures.  Needs oracle/_ref (the compiled reference).
    adeno_pp.npz            the reference's single profile-profile golden: test/adeno_fiber/
                            upgma.no_refine.part{1,2}.fasta -> upgma.pp.fasta.  Holds the two profiles'
                            score/counter tables exactly as CFAMSA::alignProfiles builds them, the rescaled gap
                            costs, and the reference outcome (traceback path, total score); generation asserts that
                            the reference run reproduces upgma.pp.fasta byte for byte.
    adeno_upgma_merges.npz  all 241 merges of test/adeno_fiber/upgma.dnd (refinement off): sequences, merge list and
                            per-merge reference outcome (total, path, CRC32 of the merged profile's scores and
                            counters as the reference's ConstructProfile built them) and the score matrix;
                            generation asserts the final alignment equals upgma.no_refine.fasta.
    """
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import pyoracle
    from treeutil import parse_newick
    T = os.path.join(REF, "adeno_fiber")
    i1, s1 = seqio.read_fasta(os.path.join(T, "upgma.no_refine.part1.fasta"))
    i2, s2 = seqio.read_fasta(os.path.join(T, "upgma.no_refine.part2.fasta"))
    ig, sg = seqio.read_fasta(os.path.join(T, "upgma.pp.fasta"))
    dp = pyoracle.RefDp(len(s1) + len(s2))            # famsa.cpp:94 adjustParams(sizes[0] + sizes[1])
    gaps = dp.gaps()
    m1, m2 = set(range(len(s1))), set(range(len(s1), len(s1) + len(s2)))
    p1 = dp.profile([s.upper() for s in s1], sorted(m1))
    p2 = dp.profile([s.upper() for s in s2], sorted(m2))
    t1, c1, k1 = dp.tables(p1)
    t2, c2, k2 = dp.tables(p2)
    m, total = dp.align(p1, p2, 1)
    rows = dp.rows(m)
    gold = {n: s.upper() for n, s in zip(ig, sg)}
    names = i1 + i2
    assert all(rows[k] == gold[names[k]] for k in rows), "reference run does not reproduce upgma.pp.fasta"
    o = pyoracle.dp_align(t1, c1, k1, t2, c2, k2, gaps)
    path = pyoracle.path_from_rows(rows, m1, m2, o["swapped"])[:len(o["path"])]
    np.savez_compressed(os.path.join(HERE, "adeno_pp.npz"), s1=t1, c1=c1, k1=k1, s2=t2, c2=c2, k2=k2, gaps=gaps,
                        path=path, total=total, swapped=o["swapped"])
    print("adeno_pp.npz: widths", t1.shape[0] - 1, t2.shape[0] - 1, "path", len(path), "total", total)
    dp.free(m); dp.close()

    ids, seqs = seqio.read_fasta(os.path.join(T, "adeno_fiber"))
    name2seq = dict(zip(ids, seqs))
    leaves, merges = parse_newick(open(os.path.join(T, "upgma.dnd")).read())
    ig, sg = seqio.read_fasta(os.path.join(T, "upgma.no_refine.fasta"))
    gold = dict(zip(ig, sg))
    lseqs = [name2seq[n].upper() for n in leaves]
    sys.path.insert(0, os.path.join(HERE, ".."))
    from dp_cases import reference_merges
    import zlib
    g, recs = reference_merges(lseqs, merges, threads=(1,), want_merged=True)
    final = recs[-1]["rows"]
    assert all(final[i] == gold[leaves[i]].upper() for i in range(len(leaves))), "does not reproduce upgma.no_refine.fasta"
    totals, paths, plen, swapped, tcrc = [], [], [], [], []
    for r in recs:
        o = pyoracle.dp_align(*r["job"], g)
        pth = pyoracle.path_from_rows(r["rows"], r["m1"], r["m2"], o["swapped"])
        totals.append(r["total"]); paths.append(pth); plen.append(len(pth)); swapped.append(o["swapped"])
        # the scores/counters ConstructProfile built for the merged profile, as CRC32 of the raw tables
        ms, mc, _ = r["merged"]
        tcrc.append((zlib.crc32(np.ascontiguousarray(ms).tobytes()), zlib.crc32(np.ascontiguousarray(mc).tobytes())))
    dp2 = pyoracle.RefDp(len(lseqs)); sm = dp2.score_matrix(); dp2.close()
    np.savez_compressed(os.path.join(HERE, "adeno_upgma_merges.npz"), seqs=np.array(lseqs), merges=np.array(merges),
                        gaps=g, totals=np.array(totals, dtype=np.int64), path=np.concatenate(paths),
                        path_len=np.array(plen), swapped=np.array(swapped), merged_crc=np.array(tcrc, dtype=np.uint32),
                        score_matrix=sm)
    print("adeno_upgma_merges.npz:", len(recs), "merges, sum path", sum(plen))


def make_hemopexin():
    """hemopexin_medoid_sl.npz -- BASELINE config 4: every merge of test/hemopexin/medoid-sl.dnd (4188 sequences,
    4187 merges, refinement is off for N > 1000).  Sequences, merge list, per-merge reference total score and a
    CRC of each traceback path; generation asserts the final alignment equals test/hemopexin/medoid-sl.fasta."""
    import zlib
    from oracle import pyoracle
    from treeutil import parse_newick
    from dp_cases import reference_merges
    T = os.path.join(REF, "hemopexin")
    ids, seqs = seqio.read_fasta(os.path.join(T, "hemopexin"))
    name2seq = dict(zip(ids, seqs))
    leaves, merges = parse_newick(open(os.path.join(T, "medoid-sl.dnd")).read())
    lseqs = [name2seq[n].upper() for n in leaves]
    g, recs = reference_merges(lseqs, merges, threads=(1,))
    ig, sg = seqio.read_fasta(os.path.join(T, "medoid-sl.fasta"))
    gold = dict(zip(ig, sg))
    final = recs[-1]["rows"]
    assert all(final[i] == gold[leaves[i]].upper() for i in range(len(leaves))), "does not reproduce medoid-sl.fasta"
    totals, crcs = [], []
    for r in recs:
        o = pyoracle.dp_align(*r["job"], g)
        pth = pyoracle.path_from_rows(r["rows"], r["m1"], r["m2"], o["swapped"])
        assert np.array_equal(pth, o["path"]) and o["total"] == r["total"]
        totals.append(r["total"]); crcs.append(zlib.crc32(pth.tobytes()))
    np.savez_compressed(os.path.join(HERE, "hemopexin_medoid_sl.npz"), seqs=np.array(lseqs), merges=np.array(merges),
                        gaps=g, totals=np.array(totals, dtype=np.int64), path_crc=np.array(crcs, dtype=np.uint32))
    print("hemopexin_medoid_sl.npz:", len(recs), "merges")


def make_hemopexin_sl():
    """hemopexin_sl.npz -- BASELINE config 4, second tree: test/hemopexin/hemopexin under the DEFAULT guide tree
    (-gt sl = MSTPrim on the length-sorted set, msa.cpp:245-256 + MSTPrim.cpp:280-549; the reference holds no golden
    file for it, so the tree is the reference's own MSTPrim run here).  Sequences in FAMSA's order, the tree's
    merges, per-merge reference total score and CRC32 of each traceback path, CRC32 of the final alignment rows."""
    import zlib
    from oracle import pyoracle
    from dp_cases import reference_merges
    ids, seqs = seqio.read_fasta(os.path.join(REF, "hemopexin", "hemopexin"))
    code_list = [seqio.encode(s.upper()) for s in seqs]
    order = sorted(range(len(seqs)), key=lambda i: (-len(code_list[i]), code_list[i].tobytes()))
    lseqs = [seqs[i].upper() for i in order]
    n = len(lseqs)
    tree = pyoracle.RefSeqSet(lseqs).mst_prim_tree(4)
    merges = [(int(a), int(b)) for a, b in tree[n:]]
    g, recs = reference_merges(lseqs, merges, threads=(1,))
    totals, crcs = [], []
    for r in recs:
        o = pyoracle.dp_align(*r["job"], g)
        pth = pyoracle.path_from_rows(r["rows"], r["m1"], r["m2"], o["swapped"])
        assert np.array_equal(pth, o["path"]) and o["total"] == r["total"]
        totals.append(r["total"]); crcs.append(zlib.crc32(pth.tobytes()))
    final = recs[-1]["rows"]
    rows_crc = zlib.crc32("\n".join(final[i] for i in range(n)).encode())
    np.savez_compressed(os.path.join(HERE, "hemopexin_sl.npz"), seqs=np.array(lseqs), merges=np.array(merges), gaps=g,
                        totals=np.array(totals, dtype=np.int64), path_crc=np.array(crcs, dtype=np.uint32),
                        rows_crc=np.array(rows_crc, dtype=np.uint32), final_width=np.array(len(final[0])))
    print("hemopexin_sl.npz:", len(recs), "merges, final width", len(final[0]))


def make_hemopexin_dups():
    """hemopexin_dups.npz -- test/hemopexin_duplicates with -keep-duplicates under its golden tree medoid-sl-dups.dnd
    (self-hosted.yml:317-329: "medoid + sl + keep duplicates (from tree)"): identical sequences are merged like any
    others.  Sequences in the tree's leaf order, merges, per-merge reference totals and path CRCs, CRC of the final rows;
    generation asserts the final alignment equals medoid-sl-dups.fasta."""
    import zlib
    from oracle import pyoracle
    from treeutil import parse_newick
    from dp_cases import reference_merges
    T = os.path.join(REF, "hemopexin_duplicates")
    ids, seqs = seqio.read_fasta(os.path.join(T, "hemopexin_duplicates"))
    name2seq = dict(zip(ids, seqs))
    leaves, merges = parse_newick(open(os.path.join(T, "medoid-sl-dups.dnd")).read())
    lseqs = [name2seq[n].upper() for n in leaves]
    g, recs = reference_merges(lseqs, merges, threads=(1,))
    ig, sg = seqio.read_fasta(os.path.join(T, "medoid-sl-dups.fasta"))
    gold = dict(zip(ig, sg))
    final = recs[-1]["rows"]
    assert all(final[i] == gold[leaves[i]].upper() for i in range(len(leaves))), "does not reproduce medoid-sl-dups.fasta"
    totals, crcs = [], []
    for r in recs:
        o = pyoracle.dp_align(*r["job"], g)
        pth = pyoracle.path_from_rows(r["rows"], r["m1"], r["m2"], o["swapped"])
        assert np.array_equal(pth, o["path"]) and o["total"] == r["total"]
        totals.append(r["total"]); crcs.append(zlib.crc32(pth.tobytes()))
    rows_crc = zlib.crc32("\n".join(final[i] for i in range(len(leaves))).encode())
    np.savez_compressed(os.path.join(HERE, "hemopexin_dups.npz"), seqs=np.array(lseqs), merges=np.array(merges), gaps=g,
                        totals=np.array(totals, dtype=np.int64), path_crc=np.array(crcs, dtype=np.uint32),
                        rows_crc=np.array(rows_crc, dtype=np.uint32))
    print("hemopexin_dups.npz:", len(recs), "merges,", len(set(lseqs)), "distinct sequences of", len(lseqs))


def make_sl_tree():
    """adeno_sl_tree.npz -- the reference's DEFAULT guide tree golden, test/adeno_fiber/sl.dnd.
    Holds the set in FAMSA's own order (length-descending, msa.cpp:245-256), the MST edges in Prim order and the
    visiting order (what famsa_lcs_prim returns) and the resulting tree_structure.  Generation asserts that
      (1) the reference's own MSTPrim run on this set,
      (2) the reference's mst_to_dendogram fed with the restated Prim's edges,
      (3) the same fed with famsa_b200.mst (Kruskal under MSTPrim's edge order + visiting-order replay)
    all give one tree, and that its clades are exactly those of sl.dnd.  Needs oracle/_ref."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from famsa_b200 import mst
    from oracle import pyoracle
    from treeutil import parse_newick
    T = os.path.join(REF, "adeno_fiber")
    ids, seqs = seqio.read_fasta(os.path.join(T, "adeno_fiber"))
    code_list = [seqio.encode(s) for s in seqs]
    order = sorted(range(len(seqs)), key=lambda i: (-len(code_list[i]), code_list[i].tobytes()))
    names = [ids[i] for i in order]
    codes, offsets, lens = seqio.pack([code_list[i] for i in order])
    n = len(lens)
    letters = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(offsets, lens)]
    ref_tree = pyoracle.RefSeqSet(letters).mst_prim_tree(2)
    # distances through the oracle, as MSTPrim's Transform<double, indel075_div_lcs> computes them
    tri = pyoracle.lcs_triangle(codes, offsets, lens)
    dist = np.array([pyoracle.transform(0, int(tri[i * (i - 1) // 2 + j]), int(lens[i]), int(lens[j]), True)
                     for i in range(1, n) for j in range(i)])
    ef, et, ed, po = mst.prim_replay(n, mst.kruskal_total_order(n, dist))
    tree = pyoracle.mst_to_dendogram(ef, et, ed, po)
    assert np.array_equal(tree, ref_tree), "edges do not rebuild the reference's MSTPrim tree"

    def clades(n_leaves, merges, leaf_names):
        sets = [frozenset([nm]) for nm in leaf_names]
        for a, b in merges:
            sets.append(sets[a] | sets[b])
        return set(sets[n_leaves:])
    gl, gm = parse_newick(open(os.path.join(T, "sl.dnd")).read())
    mine = clades(n, [tuple(int(x) for x in r) for r in tree[n:]], names)
    assert mine == clades(len(gl), gm, gl), "tree differs from the golden sl.dnd"
    np.savez_compressed(os.path.join(HERE, "adeno_sl_tree.npz"), names=np.array(names), seqs=np.array(letters),
                        edge_from=ef, edge_to=et, edge_dist=ed, prim_order=po, tree=tree)
    print("adeno_sl_tree.npz:", n, "sequences, clades equal to sl.dnd")


def clade_signatures(n_leaves, merges, leaf_hash):
    """Order-free description of a binary tree on named leaves: the multiset of (size, sum of leaf hashes mod 2^64) over
    its internal nodes -- what a set-of-clades comparison checks, in O(n) memory for 10^5 leaves."""
    size = np.ones(n_leaves + len(merges), dtype=np.int64)
    hsum = np.zeros(n_leaves + len(merges), dtype=np.uint64)
    hsum[:n_leaves] = leaf_hash
    for k, (a, b) in enumerate(merges):
        size[n_leaves + k] = size[a] + size[b]
        hsum[n_leaves + k] = hsum[a] + hsum[b]           # wraps mod 2^64
    sig = np.stack([size[n_leaves:].astype(np.uint64), hsum[n_leaves:]], axis=1)
    return sig[np.lexsort((sig[:, 1], sig[:, 0]))]


def newick_merges_iterative(text):
    """parse_newick without recursion (single-linkage trees of 10^5 leaves are 10^4 levels deep)."""
    text = text.strip().rstrip(";")
    leaves, merges, stack = [], [], []
    pos, n = 0, len(text)
    while pos < n:
        c = text[pos]
        if c == "(":
            stack.append([]); pos += 1
        elif c == ",":
            pos += 1
        elif c == ")":
            kids = stack.pop()
            pos += 1
            while pos < n and text[pos] not in ",()":
                pos += 1
            cur = kids[0]
            for k in kids[1:]:
                merges.append((cur, k)); cur = ("i", len(merges) - 1)
            if stack:
                stack[-1].append(cur)
        else:
            start = pos
            while text[pos] not in ",():":
                pos += 1
            name = text[start:pos]
            while pos < n and text[pos] not in ",()":
                pos += 1
            leaves.append(name)
            stack[-1].append(("l", len(leaves) - 1))
    nl = len(leaves)
    ident = lambda t: t[1] if t[0] == "l" else nl + t[1]
    return leaves, [(ident(a), ident(b)) for a, b in merges]


def make_lrr_sl():
    """lrr_sl.npz -- the reference's large default-guide-tree golden, test/LRR/sl.dnd (124 140 leucine-rich-repeat
    sequences, self-hosted.yml "LRR sl tree").  The fixture holds the set in FAMSA's order and the CRC32 of the
    tree_structure the reference's own MSTPrim run (oracle/_ref) builds on it; generation asserts that the clades of that
    tree are exactly those of sl.dnd.  The GPU test feeds famsa_lcs_prim's edges to the reference's mst_to_dendogram and
    must arrive at the same tree.  Needs oracle/_ref; takes a few minutes of CPU."""
    import zlib
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import pyoracle
    T = os.path.join(REF, "LRR")
    ids, seqs = seqio.read_fasta(os.path.join(T, "LRR"))
    code_list = [seqio.encode(s) for s in seqs]
    order = sorted(range(len(seqs)), key=lambda i: (-len(code_list[i]), code_list[i].tobytes()))
    names = [ids[i] for i in order]
    letters = [seqio.decode(code_list[i]) for i in order]
    n = len(letters)
    tree = pyoracle.RefSeqSet(letters).mst_prim_tree(16)
    rng = np.random.default_rng(12345)
    h = {nm: rng.integers(0, 2 ** 63, dtype=np.uint64) for nm in names}
    mine = clade_signatures(n, [tuple(int(x) for x in r) for r in tree[n:]], np.array([h[nm] for nm in names], dtype=np.uint64))
    gl, gm = newick_merges_iterative(open(os.path.join(T, "sl.dnd")).read())
    gold = clade_signatures(len(gl), gm, np.array([h[nm] for nm in gl], dtype=np.uint64))
    assert len(gl) == n and np.array_equal(mine, gold), "tree differs from the golden LRR/sl.dnd"
    crc = zlib.crc32(np.ascontiguousarray(tree[n:], dtype=np.int32).tobytes())
    np.savez_compressed(os.path.join(HERE, "lrr_sl.npz"), seqs=np.frombuffer("\n".join(letters).encode(), dtype=np.uint8),
                        tree_crc=np.array([crc], dtype=np.uint64), n=np.array([n]))
    print("lrr_sl.npz:", n, "sequences, clades equal to LRR/sl.dnd, tree crc", crc)


def make_prim_pruning_case():
    """prim_pruning_case.npz -- a small set on which MSTPrim's lower-bound skip (MSTPrim.cpp:450-467) changes the tree:
    sequences with 64-residue runs (the dropped-carry corner, LCS > shorter length) next to very short ones.  Searched with
    a fixed seed; generation asserts reference tree == restated loop with the skip != restated loop without it."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import pyoracle
    from treeutil import prim_restated
    rng = np.random.default_rng(1)
    AA = "ACDEFGHIKLMNPQRSTVWY"
    for trial in range(400):
        n = int(rng.integers(5, 14))
        seqs = []
        for _ in range(n):
            kind = rng.integers(0, 3)
            a = "ACDEFGHIKL"[rng.integers(0, 10)]
            if kind == 0:
                pre = "".join(AA[x] for x in rng.integers(0, 20, size=64 * rng.integers(0, 2)))
                s = pre + a * int(64 * rng.integers(1, 4)) + "".join(AA[x] for x in rng.integers(0, 20, size=rng.integers(0, 5)))
            else:
                s = a * int(rng.integers(1, 6)) + "".join(AA[x] for x in rng.integers(0, 20, size=rng.integers(0, 4)))
            seqs.append(s)
        seqs.sort(key=lambda q: -len(q))
        codes, off, lens = seqio.pack([seqio.encode(q) for q in seqs])
        with_skip = prim_restated(codes, off, lens, 0, True)
        without = prim_restated(codes, off, lens, 0, False)
        if all(np.array_equal(x, y) for x, y in zip(with_skip, without)):
            continue
        ref = pyoracle.RefSeqSet(seqs).mst_prim_tree(1)
        if np.array_equal(ref, pyoracle.mst_to_dendogram(*without)):
            continue                                    # the edge lists differ but give the same tree: not a witness
        assert np.array_equal(ref, pyoracle.mst_to_dendogram(*with_skip)), "restated loop with the skip differs from the reference"
        np.savez_compressed(os.path.join(HERE, "prim_pruning_case.npz"), seqs=np.array(seqs), edge_from=with_skip[0], edge_to=with_skip[1],
                            edge_dist=with_skip[2], prim_order=with_skip[3], tree=ref)
        print("prim_pruning_case.npz: trial", trial, n, "sequences; reference == loop with the skip, != loop without")
        return
    raise SystemExit("no witness found")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.path.insert(0, os.path.join(HERE, ".."))
        globals()[sys.argv[1]]()
        sys.exit(0)
    main()
    make_dp()
    make_hemopexin()
    make_hemopexin_sl()
    make_hemopexin_dups()
    make_sl_tree()
    make_prim_pruning_case()
