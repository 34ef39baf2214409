"""One pass of bench.py's DP workload through the host-buffer ABI (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, famsa_b200
rows, jobs = bench.dp_workload(0)
eng = famsa_b200.Engine(0)
g = np.array(bench.DP_GAPS, dtype=np.int64)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    eng.dp_align_batch(jobs, g)
print(eng.dp_last_timing())
