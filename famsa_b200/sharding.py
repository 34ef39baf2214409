"""Multi-GPU plumbing for HP-1: the triangle shards by rows, one process per GPU, and the only exchange step is
one all-gather of the row blocks (NCCL on GPUs, gloo in the CPU tests).  No compute lives here."""
from __future__ import annotations

import math


def tri(r: int) -> int:
    return r * (r - 1) // 2 if r else 0


def row_shards(n: int, parts: int, align: int = 32) -> list[int]:
    """Row boundaries b[0]=0 <= ... <= b[parts]=n giving every rank (almost) the same number of pairs
    (row i holds i pairs, so boundaries sit at n*sqrt(r/parts)), rounded to the 32-row mask groups of the LCS kernel so
    that no group's tiles are computed by two ranks."""
    total = tri(n)
    bounds = [0]
    for r in range(1, parts):
        target = total * r / parts
        b = int(round((1 + math.sqrt(1 + 8 * target)) / 2))
        if align > 1 and n >= 4 * align * parts:
            b = (b + align // 2) // align * align
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds


def shard_sizes(bounds: list[int]) -> list[int]:
    return [tri(bounds[r + 1]) - tri(bounds[r]) for r in range(len(bounds) - 1)]


def all_gather_blocks(block, bounds, dist, out=None):
    """block: 1-D tensor holding this rank's packed rows [bounds[rank], bounds[rank+1]) padded to the largest
    shard.  Returns the gathered (world * max_shard) tensor; full_triangle() strips the padding."""
    import torch
    world = len(bounds) - 1
    if out is None:
        out = torch.empty(block.numel() * world, dtype=block.dtype, device=block.device)
    # gather raw bytes: every backend (gloo in the CPU tests, NCCL on GPUs) moves uint8
    dist.all_gather_into_tensor(out.view(torch.uint8), block.view(torch.uint8))
    return out


def full_triangle(gathered, bounds):
    """Concatenate the real part of every rank's block -> the packed lower triangle of all n rows."""
    import torch
    sizes = shard_sizes(bounds)
    m = max(sizes)
    return torch.cat([gathered[r * m:r * m + sizes[r]] for r in range(len(sizes))])


def sub_bounds(row_begin: int, row_end: int, n_sub: int) -> list[int]:
    """Splits rows [row_begin, row_end) into n_sub consecutive pieces holding (almost) the same number of pairs."""
    lo, hi = tri(row_begin), tri(row_end)
    b = [row_begin]
    for k in range(1, n_sub):
        target = lo + (hi - lo) * k / n_sub
        r = int(round((1 + math.sqrt(1 + 8 * target)) / 2))
        b.append(min(max(r, b[-1]), row_end))
    b.append(row_end)
    return b


_side_stream = None


def triangle_allgather_pipelined(compute_rows, bounds, rank, dist, full, n_sub: int = 4):
    """The N>1 triangle with its exchange step overlapped: every rank computes its row shard in n_sub pieces, and as
    soon as a piece is finished every rank's piece of that number is broadcast straight into its place in `full` (the
    packed lower triangle of all n rows, identical on every rank afterwards) while the next piece is being computed --
    no padding, no staging copy, and only the last piece's exchange is exposed.
    compute_rows(r0, r1, view): queues the computation of rows [r0, r1) on the current stream, writing the packed rows
    into `view` (= full[tri(r0):tri(r1)]).  Works on CPU tensors too (gloo; everything is then synchronous)."""
    import torch
    global _side_stream
    world = len(bounds) - 1
    subs = [sub_bounds(bounds[r], bounds[r + 1], n_sub) for r in range(world)]
    cuda = full.is_cuda
    works = []
    if cuda:
        main = torch.cuda.current_stream()
        if _side_stream is None:
            _side_stream = torch.cuda.Stream()
    for b in range(n_sub):
        r0, r1 = subs[rank][b], subs[rank][b + 1]
        if r1 > r0 and tri(r1) > tri(r0):
            compute_rows(r0, r1, full[tri(r0):tri(r1)])
        # raw bytes: every backend (gloo in the CPU tests, NCCL on GPUs) moves uint8
        views = [(src, full[tri(subs[src][b]):tri(subs[src][b + 1])].view(torch.uint8)) for src in range(world)]
        if cuda:
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(_side_stream):
                _side_stream.wait_event(ev)
                for src, v in views:
                    if v.numel():
                        works.append(dist.broadcast(v, src=src, async_op=True))
        else:
            for src, v in views:
                if v.numel():
                    dist.broadcast(v, src=src)
    for w in works:
        w.wait()                      # the current stream waits for the exchange; the host does not
    if cuda:
        main.wait_stream(_side_stream)
    return full


class PeerTriangle:
    """The N>1 triangle with the exchange folded into the computation (famsa_lcs_triangle_exchange): every rank owns a
    device buffer for the full packed triangle, exported through CUDA IPC and mapped by every other rank, so a finished
    piece of a rank's rows is copied straight into every peer's buffer over NVLink by the copy engines while the next
    piece is computed -- no collective kernel, no staging, no padding.  `dist` only carries the 64-byte handles once and
    the barrier that ends a step."""

    def __init__(self, eng, n: int, elem_bytes: int, rank: int, world: int, dist):
        self.eng, self.n, self.elem_bytes, self.rank, self.world, self.dist = eng, n, elem_bytes, rank, world, dist
        self.nbytes = max(tri(n) * elem_bytes, 1)
        self.ptr = eng.device_alloc(self.nbytes)
        handles = [None] * world
        dist.all_gather_object(handles, eng.ipc_export(self.ptr))
        self.peers = [eng.ipc_open(handles[r]) for r in range(world) if r != rank]
        self.bounds = row_shards(n, world)

    def tensor(self, torch):
        """The local buffer as a torch tensor (zero copy)."""
        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (max(tri(self.n), 1),), "typestr": "<i2" if self.elem_bytes == 2 else "<i4",
                                      "data": (self.ptr, False), "version": 2}
        return torch.as_tensor(a, device="cuda")[:tri(self.n)]

    def step(self, torch, stream: int = 0, n_pieces: int = 8, flag=None):
        """Queues this rank's rows + their copies on `stream`, then the barrier: with a CUDA `flag` tensor a one-element
        all-reduce ordered on the current stream (NCCL), else a host barrier after a device synchronise (gloo)."""
        rb, re = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.eng.triangle_exchange(rb, re, self.ptr, self.peers, self.elem_bytes, n_pieces, stream)
        if flag is not None and flag.is_cuda:
            self.dist.all_reduce(flag)
        else:
            torch.cuda.synchronize()
            self.dist.barrier()

    def close(self, torch):
        torch.cuda.synchronize()
        self.dist.barrier()                        # nobody unmaps a buffer a peer may still be writing to
        for p in self.peers:
            self.eng.ipc_close(p)
        self.dist.barrier()
        self.eng.device_free(self.ptr)
        self.peers, self.ptr = [], 0


def group_slice(n_groups: int, shard: int, n_shards: int) -> tuple[int, int]:
    """The run of 32-sequence mask groups (length-sorted order) rank `shard` answers for -- same split as
    famsa_lcs_assign_shard."""
    return n_groups * shard // n_shards, n_groups * (shard + 1) // n_shards


def assign_allreduce(assign_shard, packed, rank: int, world: int, dist):
    """Medoid assignment (FastTree<>::makeEvaluation, FastTree.cpp:309-324) over `world` ranks: rank r computes the seed
    rows against its slice of the sequences only (assign_shard(r, world, packed) fills the int64 tensor `packed`, see
    famsa_lcs_assign_shard), then ONE element-wise MIN all-reduce completes the (distance, seed) pairs on every rank.
    Returns `packed` (reduced in place); binding.unpack_assignment() splits it."""
    assign_shard(rank, world, packed)
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MIN)
    return packed
