"""GPU test of famsa_lcs_triangle_exchange (the N>1 triangle: rows sharded over ranks, every finished piece copied into
the peers' full-triangle buffers through CUDA IPC while the next piece is computed).  Two processes, one rank each; on a
box with a single GPU both ranks share device 0 (IPC between processes works on one device too), the 64-byte handles and
the barrier travel over gloo.  Every rank must end with the whole packed triangle, equal to the single-context one."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, length, n_pieces, sort_desc, out_dir, sorted_space=False):
    import torch
    import torch.distributed as dist
    from famsa_b200 import seqio, sharding
    from famsa_b200.binding import Engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = Engine(dev)
    codes, offsets, lens = seqio.synth_family(n, length, seed=9, sort_desc=sort_desc)
    order = eng.upload_sorted(codes, offsets, lens) if sorted_space else None
    if not sorted_space:
        eng.upload(codes, offsets, lens)
    pt = sharding.PeerTriangle(eng, n, 2, rank, world, dist)
    full = pt.tensor(torch)
    ok = True
    for rep in range(2):
        full.fill_(-1)
        torch.cuda.synchronize(); dist.barrier()
        pt.step(torch, 0, n_pieces)
        tiles = eng.last_tiles()
        got = full.cpu().numpy().astype(np.uint16)
        want = eng.triangle(0, n, dtype=np.uint16)
        ok = ok and bool(np.array_equal(got, want))
        if sorted_space:
            # a rank launches only its share of the tiles (+ at most the mask groups its boundaries cut) ...
            ok = ok and tiles <= eng.last_tiles() / world * 1.05 + 4 * (n // 64 + 1)
            # ... and element (i, j) of the caller's order is element (pos[i], pos[j]) of the gathered triangle
            ref = Engine(dev); ref.upload(codes, offsets, lens)
            caller = ref.triangle(0, n, dtype=np.uint16); ref.close()
            pos = np.empty(n, dtype=np.int64); pos[order] = np.arange(n)
            i, j = np.tril_indices(n, -1)
            a, b = np.maximum(pos[i], pos[j]), np.minimum(pos[i], pos[j])
            ok = ok and bool(np.array_equal(caller[i * (i - 1) // 2 + j], got[a * (a - 1) // 2 + b]))
    pt.close(torch)
    eng.close()
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,length,n_pieces,sort_desc", [(2, 1500, 120, 4, True), (3, 700, 90, 8, True), (2, 333, 70, 1, True),
                                                               (2, 400, 80, 4, False)])
def test_triangle_exchange_over_ipc(tmp_path, world, n, length, n_pieces, sort_desc):
    import torch.multiprocessing as mp
    port = 29800 + world * 11 + (os.getpid() % 60)
    mp.spawn(_worker, args=(world, port, n, length, n_pieces, sort_desc, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_triangle_exchange_unsorted_input_sorted_space(tmp_path):
    """famsa_lcs_upload_sorted: an unsorted set sharded in the library's own order -- every rank launches 1/N of the tiles
    (checked through famsa_lcs_last_tiles) and the gathered triangle, read through the returned permutation, is the
    caller-order triangle."""
    import torch.multiprocessing as mp
    world, port = 2, 29900 + (os.getpid() % 60)
    mp.spawn(_worker, args=(world, port, 900, 110, 4, False, str(tmp_path), True), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
