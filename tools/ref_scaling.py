"""Thread scaling of the reference CPU path on this host (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from famsa_b200 import seqio
from oracle import pyoracle
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(p):
        print(p, open(p).read().strip())
codes, offsets, lens = seqio.synth_family(10000, 400, 1)
letters = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(offsets, lens)]
rs = pyoracle.RefSeqSet(letters)
for t in (8, 16, 32, 64, 128):
    sec, pairs, _ = rs.triangle_mt(7000, 10000, t, 2)
    print(f"threads {t:4d}: {pairs/sec/1e6:8.2f} Mpairs/s  ({sec:.2f} s)")
