"""Where the exposed part of the N>1 triangle step goes: torchrun --nproc-per-node N tools/exchange_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench, famsa_b200
from famsa_b200.sharding import PeerTriangle, tri

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
(codes, offsets, lens), n = bench.workload(world)
eng = famsa_b200.Engine(local)
eng.upload(codes, offsets, lens)
pt = PeerTriangle(eng, n, 2, rank, world, dist)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); stream = side.cuda_stream
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
rb, re = pt.bounds[rank], pt.bounds[rank + 1]

def timed(fn, reps=4):
    for _ in range(2): fn()
    dist.barrier(); torch.cuda.synchronize()
    tot = 0.0
    for r in range(reps):
        flush.fill_(r)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    t = torch.tensor([tot / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = {}
res["whole block, one launch (triangle_device)"] = timed(lambda: eng.triangle_device(rb, re, pt.ptr + 2 * tri(rb), 2, stream))
for np_ in (1, 2, 4, 8):
    res[f"exchange pieces={np_}, no peers, no barrier"] = timed(lambda: eng.triangle_exchange(rb, re, pt.ptr, [], 2, np_, stream))
    res[f"exchange pieces={np_}, peers, no barrier"] = timed(lambda: eng.triangle_exchange(rb, re, pt.ptr, pt.peers, 2, np_, stream))
    res[f"exchange pieces={np_}, peers + all-reduce"] = timed(lambda: pt.step(torch, stream, np_, flag))
res["all-reduce alone"] = timed(lambda: dist.all_reduce(flag))
if rank == 0:
    for k, v in res.items():
        print(f"{v:9.3f} ms  {k}")
pt.close(torch)
dist.destroy_process_group()
