"""world_size-2 (and 3) gloo test of the N>1 path's host logic: equal-pair row shards + one all-gather of the
row blocks reproduce the full packed triangle.  The per-rank compute is the oracle here (no GPU in this tier)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from famsa_b200 import seqio, sharding
from oracle import pyoracle


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    codes, offsets, lens = seqio.synth_family(n, 60, seed=4)
    bounds = sharding.row_shards(n, world)
    rb, re = bounds[rank], bounds[rank + 1]
    sizes = sharding.shard_sizes(bounds)
    block = torch.zeros(max(sizes), dtype=torch.int16)
    mine = pyoracle.lcs_triangle(codes, offsets, lens, rb, re).astype(np.int16)
    block[:mine.size] = torch.from_numpy(mine)
    gathered = sharding.all_gather_blocks(block, bounds, dist)
    full = sharding.full_triangle(gathered, bounds).numpy().astype(np.uint32)
    want = pyoracle.lcs_triangle(codes, offsets, lens)
    ok = np.array_equal(full, want) and sum(sizes) == want.size
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 101), (3, 64)])
def test_row_shards_all_gather(tmp_path, world, n):
    port = 29600 + world * 7 + (os.getpid() % 50)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_row_shards_balance():
    for n, parts in [(10000, 2), (20000, 4), (28284, 8), (7, 8), (1, 2)]:
        b = sharding.row_shards(n, parts)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        sizes = sharding.shard_sizes(b)
        assert sum(sizes) == n * (n - 1) // 2
        if n >= 1000:
            assert max(sizes) <= 1.01 * (sum(sizes) / parts)


def test_schedule_levels_and_shards():
    from famsa_b200.schedule import ready_levels, shard_level
    merges = [(0, 1), (2, 3), (4, 5), (6, 7), (8, 9)]        # 6 leaves: (0,1)->6 (2,3)->7 (4,5)->8 (6,7)->9 (8,9)->10
    lv = ready_levels(6, merges)
    assert lv == [[0, 1, 2], [3], [4]]
    parts = shard_level([0, 1, 2, 3], [10, 7, 5, 4], 2)
    assert sorted(sum(parts, [])) == [0, 1, 2, 3] and abs(sum([10, 7, 5, 4][k] for k in parts[0]) - 13) <= 3
