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
