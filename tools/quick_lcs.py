"""Quick device-side timing of the LCS triangle (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import famsa_b200
from famsa_b200 import seqio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
codes, offsets, lens = seqio.synth_family(n, L, seed=1)
eng = famsa_b200.Engine(0)
t = time.time(); eng.upload(codes, offsets, lens); print("upload s", time.time() - t)
pairs = n * (n - 1) // 2
d_out = torch.empty(pairs, dtype=torch.int16, device="cuda")
for r in range(reps):
    t = time.time()
    eng.triangle_device(0, n, d_out.data_ptr(), 2)
    wall = time.time() - t
    tot, main, p = eng.last_timing()
    print(f"rep {r}: wall {wall*1e3:.1f} ms  total {tot:.2f} ms  main {main:.2f} ms  -> {p/main/1e6:.1f} Mpairs/s (main), {p/tot/1e6:.1f} (total)")
print("checksum", int(d_out.to(torch.int64).sum().item()))
