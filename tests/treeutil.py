"""Tiny Newick reader and merge-order helpers for the tests (guide trees are inputs to HP-2, the
tree builders themselves are out of scope)."""
from __future__ import annotations


def parse_newick(text: str):
    """Returns (leaf_names, merges): merges is a list of (left, right) node ids in post-order; leaves are
    0..n-1 in order of appearance, internal node k gets id n+k -- the layout of the reference's
    tree_structure (src/tree/TreeDefs.h:15-16)."""
    text = text.strip().rstrip(";")
    pos = 0
    leaves: list[str] = []
    raw_merges: list[tuple] = []

    def node():
        nonlocal pos
        if text[pos] == "(":
            pos += 1
            kids = [node()]
            while text[pos] == ",":
                pos += 1
                kids.append(node())
            assert text[pos] == ")"
            pos += 1
            skip_label()
            cur = kids[0]
            for k in kids[1:]:
                raw_merges.append((cur, k))
                cur = ("i", len(raw_merges) - 1)
            return cur
        start = pos
        while text[pos] not in ",():":
            pos += 1
        name = text[start:pos]
        skip_label()
        leaves.append(name)
        return ("l", len(leaves) - 1)

    def skip_label():
        nonlocal pos
        while pos < len(text) and text[pos] not in ",()":
            pos += 1

    node()
    n = len(leaves)
    ident = lambda t: t[1] if t[0] == "l" else n + t[1]
    return leaves, [(ident(a), ident(b)) for a, b in raw_merges]


def levels(n_leaves: int, merges):
    from famsa_b200.schedule import ready_levels
    return ready_levels(n_leaves, merges)


def prim_restated(codes, offsets, lens, kind, prune=True):
    """MSTPrim<>::run_view's vertex loop (MSTPrim.cpp:280-549) restated with the oracle, including the lower-bound skip of
    :450-467 (a candidate is only computed when the distance it would have with LCS = the shorter length does not exceed
    its current one; without the dropped-carry corner that never changes a result).  Validated against the reference's own
    tree in test_gpu_prim_tree and at the generation of prim_pruning_case.npz."""
    import numpy as np
    from oracle import pyoracle
    n = len(lens)
    lcs = pyoracle.lcs_rows(codes, offsets, lens, np.arange(n))          # lcs[v][j], v = row (seq0)
    dist = np.full(n, np.finfo(np.float64).max)
    key = np.zeros(n, dtype=np.uint64)
    visited = np.zeros(n, dtype=bool)
    order = np.full(n, n, dtype=np.int32)
    full = np.uint64(0xFFFFFFFFFFFFFFFF)
    v = 0
    visited[0] = True
    order[0] = 0
    ef, et, ed = [], [], []
    for step in range(1, n):
        best = -1
        for j in range(n):
            if visited[j]:
                continue
            if not prune or pyoracle.transform(kind, int(min(lens[v], lens[j])), int(lens[v]), int(lens[j]), True) <= dist[j]:
                d = pyoracle.transform(kind, int(lcs[v, j]), int(lens[v]), int(lens[j]), True)
                if d <= dist[j]:
                    a, b = (v, j) if v < j else (j, v)
                    k = full ^ np.uint64((a << 32) + b)
                    if d < dist[j] or k < key[j]:
                        dist[j], key[j] = d, k
            if best < 0 or dist[j] < dist[best] or (dist[j] == dist[best] and key[j] < key[best]):
                best = j
        p = int(full ^ key[best])
        a, b = p >> 32, p & 0xFFFFFFFF
        ef.append(min(a, b)); et.append(max(a, b)); ed.append(dist[best])
        order[best] = step
        visited[best] = True
        v = best
    return np.array(ef, np.int32), np.array(et, np.int32), np.array(ed), order
