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


def run_tree(eng):
    """The whole progressive alignment behind test/adeno_fiber/upgma.no_refine.fasta in one famsa_prof_align_tree call:
    241 merges, totals and paths as the reference produced them (fixture adeno_upgma_merges.npz)."""
    from famsa_b200 import seqio
    z = np.load(os.path.join(GOLDEN, "adeno_upgma_merges.npz"))
    seqs = [str(s) for s in z["seqs"]]
    codes, off, lens = seqio.pack([seqio.encode(s) for s in seqs])
    eng.upload(codes, off, lens)
    eng.prof_set_scoring(z["score_matrix"])
    root, res, st = eng.align_tree([tuple(int(x) for x in m) for m in z["merges"]], z["gaps"])
    assert [r["total"] for r in res] == [int(t) for t in z["totals"]]
    assert np.array_equal(np.concatenate([r["path"] for r in res]), z["path"])
    eng.prof_drop([root])
    print(f"smoke ok: {len(res)} merges of the adeno_fiber UPGMA tree in one call, {st['cells']} cells, {st['wall_ms']:.2f} ms, bit-exact")
