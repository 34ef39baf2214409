"""Cost of one wavefront step and of the lag between stripes: single merges of growing height, one at a time.
usage: step_cost.py [pp|sp]   (run under ncu --metrics gpu__time_duration.sum to get the fill kernel alone)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, famsa_b200
from famsa_b200 import profiles
which = sys.argv[1] if len(sys.argv) > 1 else "pp"
rng = np.random.default_rng(3)
sm = profiles.synth_score_matrix(rng)
eng = famsa_b200.Engine(0)
g = np.array(bench.DP_GAPS, dtype=np.int64)
WC = 4000
for rows in (31, 63, 127, 255, 511, 991):
    a = profiles.tables_from_rows(profiles.synth_alignment(40 if which == "pp" else 1, rows, rng), sm, bench.DP_GAPS)
    b = profiles.tables_from_rows(profiles.synth_alignment(40, WC, rng), sm, bench.DP_GAPS)
    job = [(a[0], a[1], a[2], b[0], b[1], b[2])]
    eng.dp_align_batch(job, g)
    r = eng.dp_align_batch(job, g)
    t = eng.dp_last_timing()
    print(which, rows, WC, "swapped", r[0]["swapped"], "variant", r[0].get("variant"), "timing", t, flush=True)
