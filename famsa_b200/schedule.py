"""Level-synchronous scheduling of guide-tree merges -- the batch-shaped replacement for the reference's
CProfileQueue (src/core/queues.cpp:17-187), which hands ready merges to worker threads one at a time
(deepest-ready-first).  On a GPU the unit of submission is "every merge whose children are finished", so the host
loop becomes:  for level in ready_levels(tree): famsa_dp_align_batch(level) ; ConstructProfile for each (host).

tree: the reference's tree_structure (src/tree/TreeDefs.h:15-16) -- n leaves (ids 0..n-1) followed by the internal
nodes; merges[k] = (left, right) child ids of internal node n+k, children always precede parents.
"""
from __future__ import annotations


def ready_levels(n_leaves: int, merges) -> list[list[int]]:
    """Group merge indices into dependency levels: level 0 merges only leaves, level d has a child of level d-1."""
    depth = [0] * (n_leaves + len(merges))
    out: list[list[int]] = []
    for k, (a, b) in enumerate(merges):
        d = max(depth[a], depth[b]) + 1
        depth[n_leaves + k] = d
        while len(out) < d:
            out.append([])
        out[d - 1].append(k)
    return out


def shard_level(level: list[int], costs: list[int], n_ranks: int) -> list[list[int]]:
    """Greedy longest-processing-time split of one level's merges over ranks (multi-GPU: merges of a level are
    independent, a single merge does not shard).  costs[k] = W1*W2 of merge k."""
    loads = [0] * n_ranks
    parts: list[list[int]] = [[] for _ in range(n_ranks)]
    for k in sorted(level, key=lambda k: -costs[k]):
        r = loads.index(min(loads))
        parts[r].append(k)
        loads[r] += costs[k]
    return parts


# ------------------------------------------------------------------------------------------------------------------
# Multi-GPU progressive alignment with resident profiles: shard by SUBTREE.
#
# Merges of one level are independent, but with profiles kept in HBM (famsa_prof_merge_batch) a merge wants both
# children on the same device.  Whole subtrees have that property by construction, so the guide tree is cut into a
# frontier of disjoint subtrees that are dealt to the ranks; every rank aligns its subtrees with no communication at
# all, then the few subtree roots travel to rank 0 (384 B per column, once) and rank 0 finishes the top of the tree.
# A single merge still does not shard ("replicas only" per merge).

def subtree_shards(n_leaves: int, merges, world: int, oversplit: int = 4):
    """Returns (owner, frontier): owner[k] = rank that executes merge k, or -1 for the top merges that rank 0 runs after
    the hand-over; frontier = list of (node id, rank) of the subtree roots (a leaf can be a frontier node).
    The cost of a subtree is its number of merges (each costs about W1*W2 = L^2 cells whatever the cardinality)."""
    n_nodes = n_leaves + len(merges)
    size = [0] * n_nodes                              # merges inside the subtree rooted at the node
    for k, (a, b) in enumerate(merges):
        size[n_leaves + k] = size[a] + size[b] + 1
    owner = [-1] * len(merges)
    if not merges:
        return owner, []
    root = n_nodes - 1
    frontier = [root]
    target = max(1, world * oversplit)
    # split the largest frontier subtree until there are enough pieces to balance (its own merge joins the top)
    while len(frontier) < target:
        big = max(frontier, key=lambda v: size[v])
        if size[big] == 0:
            break
        frontier.remove(big)
        frontier.extend(merges[big - n_leaves])
    loads = [0] * world
    assign = []
    for v in sorted(frontier, key=lambda v: -size[v]):
        r = loads.index(min(loads))
        loads[r] += size[v]
        assign.append((v, r))
    node_rank = dict(assign)
    # every merge below a frontier node inherits that node's rank (children precede parents: walk downwards)
    for k in range(len(merges) - 1, -1, -1):
        v = n_leaves + k
        if v in node_rank:
            owner[k] = node_rank[v]
            for c in merges[k]:
                if c >= n_leaves:
                    node_rank[c] = node_rank[v]
    return owner, assign


def sharded_resident_alignment(engine, dist, rank: int, world: int, n_leaves: int, leaf_widths, merges, gaps):
    """Runs the whole guide tree on `world` ranks (one engine = one GPU per rank): own subtrees first, level by level,
    with no communication; then the subtree roots are handed to rank 0 (famsa_prof_get -> all_gather_object ->
    famsa_prof_put), which runs the top merges.  `engine` needs the sequences uploaded and the scoring set.
    Returns {merge index: result dict} for the merges this rank executed (rank 0: including the top) and the
    resident id of the root on rank 0 (None elsewhere).  dist: torch.distributed (any backend with object collectives)
    or None for world == 1."""
    from .binding import PROF_LEAF
    owner, frontier = subtree_shards(n_leaves, merges, world)
    node = {}                                          # node id -> resident id / leaf handle on this rank
    width = {i: int(leaf_widths[i]) for i in range(n_leaves)}
    results = {}

    def handle(v):
        return node[v] if v in node else PROF_LEAF | v

    def run(indices):
        for lvl in ready_levels_subset(n_leaves, merges, indices):
            pairs = [(handle(merges[k][0]), handle(merges[k][1])) for k in lvl]
            ids, res = engine.prof_merge_batch(pairs, gaps, [(width[merges[k][0]], width[merges[k][1]]) for k in lvl])
            for k, pid, r in zip(lvl, ids, res):
                node.pop(merges[k][0], None); node.pop(merges[k][1], None)
                node[n_leaves + k] = pid
                width[n_leaves + k] = len(r["path"])
                results[k] = r

    run([k for k in range(len(merges)) if owner[k] == rank])
    # hand the subtree roots over to rank 0
    mine = [(v, engine.prof_get(node[v]) + (width[v],)) for v, r in frontier if r == rank and v >= n_leaves and rank != 0]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    if rank != 0:
        if mine:
            engine.prof_drop([node[v] for v, _ in mine])
        return results, None
    for part in gathered[1:]:
        if part:
            ids = engine.prof_put([(s, c, card) for _, (s, c, card, _) in part])
            for (v, (_, _, _, w)), pid in zip(part, ids):
                node[v] = pid
                width[v] = w
    # (leaves that are frontier nodes of other ranks need nothing: rank 0 materialises any leaf itself)
    run([k for k in range(len(merges)) if owner[k] == -1])
    return results, node[n_leaves + len(merges) - 1]


def ready_levels_subset(n_leaves: int, merges, indices) -> list[list[int]]:
    """ready_levels restricted to a subset of merges: children outside the subset count as finished."""
    inside = set(indices)
    depth = {}
    out: list[list[int]] = []
    for k in sorted(inside):
        a, b = merges[k]
        d = 1 + max(depth.get(a, 0), depth.get(b, 0))
        depth[n_leaves + k] = d
        while len(out) < d:
            out.append([])
        out[d - 1].append(k)
    return out
