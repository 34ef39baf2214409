"""smoke(): one small HP-2 batch on cuda:0 checked against the oracle (fixture adeno_pp.npz)."""
import os

import numpy as np

from oracle import pyoracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(eng):
    z = np.load(os.path.join(GOLDEN, "adeno_pp.npz"))
    job = (z["s1"], z["c1"], int(z["k1"]), z["s2"], z["c2"], int(z["k2"]))
    got = eng.dp_align_batch([job, job], z["gaps"], want_dirs=True)
    want = pyoracle.dp_align(*job, z["gaps"])
    for g in got:
        assert g["total"] == want["total"] == int(z["total"])
        assert np.array_equal(g["path"], want["path"]) and np.array_equal(g["path"], z["path"])
        assert np.array_equal(g["dirs"], want["dirs"])
    print(f"smoke ok: profile-profile DP {want['dirs'].shape} bit-exact (score {want['total']}, path {len(want['path'])})")
