"""One very wide merge (top of the guide tree) through famsa_prof_merge_batch.  usage: big_merge.py W1 W2 card1 card2 [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import famsa_b200
from famsa_b200 import profiles

w1, w2, k1, k2 = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(3)
sm = profiles.synth_score_matrix(rng)
gaps = np.array((-14850, -1250, -660, -660), dtype=np.int64)
a = profiles.tables_from_rows(profiles.synth_alignment(k1, w1, rng), sm, gaps)
b = profiles.tables_from_rows(profiles.synth_alignment(k2, w2, rng), sm, gaps)
eng = famsa_b200.Engine(0)
for r in range(reps):
    ids = eng.prof_put([a, b])
    t = time.time(); m, res = eng.prof_merge_batch([(ids[0], ids[1])], gaps, [(w1, w2)]); call = time.time() - t
    tot, con = eng.prof_last_timing(); _, k, c = eng.dp_last_timing()
    print(f"rep {r}: cells {c} swapped {res[0]['swapped']} path {len(res[0]['path'])} dp_kernels {k:.3f} ms device {tot:.3f} ms "
          f"construct {con:.3f} ms call {call * 1e3:.3f} ms  {c / k / 1e6:.3f} Gcells/s")
    eng.prof_drop(m)
