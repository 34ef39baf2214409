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
