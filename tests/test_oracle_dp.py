"""CPU tests: oracle/dp_oracle.c against the reference's golden alignments and the reference itself."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from dp_cases import check_against_reference, random_tree, reference_merges
from famsa_b200 import seqio
from oracle import pyoracle

needs_ref = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built")


def test_oracle_pp_golden():
    """test/adeno_fiber/upgma.pp.fasta: the reference's one pure profile-profile known answer."""
    z = np.load(os.path.join(GOLDEN, "adeno_pp.npz"))
    o = pyoracle.dp_align(z["s1"], z["c1"], int(z["k1"]), z["s2"], z["c2"], int(z["k2"]), z["gaps"])
    assert o["variant"] == 2 and o["swapped"] == bool(z["swapped"])
    assert o["total"] == int(z["total"])
    assert np.array_equal(o["path"], z["path"])
    d = o["dirs"]
    assert np.all(d[0, 1:] == 0x15) and np.all(d[1:, 0] == 0x2A) and d[0, 0] == 0


@needs_ref
def test_oracle_all_merges_of_golden_upgma_tree():
    """All 241 merges (SeqSeq, SeqProf, ProfProf) behind test/adeno_fiber/upgma.no_refine.fasta."""
    z = np.load(os.path.join(GOLDEN, "adeno_upgma_merges.npz"))
    seqs = [str(s) for s in z["seqs"]]
    merges = [tuple(int(x) for x in m) for m in z["merges"]]
    g, recs = reference_merges(seqs, merges, threads=(1, 2))
    assert np.array_equal(g, z["gaps"])
    res = [pyoracle.dp_align(*r["job"], g) for r in recs]
    check_against_reference(res, recs)
    assert [r["total"] for r in res] == [int(t) for t in z["totals"]]
    assert np.array_equal(np.concatenate([r["path"] for r in res]), z["path"])
    assert sorted(set(r["variant"] for r in res)) == [0, 1, 2]


@needs_ref
@pytest.mark.parametrize("seed,n,length,gaps", [(1, 60, 90, None), (2, 40, 300, None), (3, 50, 40, (-9000, -700, -300, -100)),
                                                (4, 30, 150, (-20000, -2000, -2500, -900))])
def test_oracle_random_families(seed, n, length, gaps):
    rng = np.random.default_rng(seed)
    codes, off, lens = seqio.synth_family(n, length, seed, sort_desc=False)
    seqs = [seqio.decode(codes[int(o):int(o) + int(l)]) for o, l in zip(off, lens)]
    seqs[0] = seqs[0][:5] + "XBZ" + seqs[0][8:]          # non-standard residues
    merges = random_tree(n, rng)
    g, recs = reference_merges(seqs, merges, threads=(1, 2), rng=rng, gaps=gaps)
    res = [pyoracle.dp_align(*r["job"], g) for r in recs]
    check_against_reference(res, recs)


def test_oracle_hemopexin_fixture_selfcheck():
    z = np.load(os.path.join(GOLDEN, "hemopexin_medoid_sl.npz"))
    assert len(z["seqs"]) == 4188 and len(z["merges"]) == 4187 and len(z["totals"]) == 4187


@needs_ref
def test_oracle_hemopexin_first_levels():
    """Config 4 subset on CPU (the whole tree runs in the GPU test): first 400 merges of medoid-sl.dnd."""
    z = np.load(os.path.join(GOLDEN, "hemopexin_medoid_sl.npz"))
    seqs = [str(s) for s in z["seqs"]]
    merges = [tuple(int(x) for x in m) for m in z["merges"]]
    # keep a prefix that is closed under dependencies
    n = len(seqs)
    keep = []
    for k, (a, b) in enumerate(merges):
        if len(keep) >= 400:
            break
        if (a < n or (a - n) in keep) and (b < n or (b - n) in keep):
            keep.append(k)
    remap = {k: i for i, k in enumerate(keep)}
    sub = [tuple(x if x < n else n + remap[x - n] for x in merges[k]) for k in keep]
    g, recs = reference_merges(seqs, sub, n_seqs_for_rescale=n, threads=(1,))
    assert np.array_equal(g, z["gaps"])
    for k, r in zip(keep, recs):
        o = pyoracle.dp_align(*r["job"], g)
        assert o["total"] == int(z["totals"][k])
        assert zlib.crc32(o["path"].tobytes()) == int(z["path_crc"][k])


@needs_ref
def test_oracle_driven_alignment_equals_reference():
    """The oracle's direction matrices drive the reference's own ConstructProfile through the whole upgma tree of
    adeno_fiber: the final alignment must be the reference's, row for row."""
    from dp_cases import driven_progressive_alignment
    z = np.load(os.path.join(GOLDEN, "adeno_upgma_merges.npz"))
    seqs = [str(s) for s in z["seqs"]]
    merges = [tuple(int(x) for x in m) for m in z["merges"]]
    rows, total = driven_progressive_alignment(
        seqs, merges, lambda jobs, g: [pyoracle.dp_align(*j, g) for j in jobs])
    _, recs = reference_merges(seqs, merges, threads=(1,))
    assert rows == recs[-1]["rows"] and total == recs[-1]["total"] == int(z["totals"][-1])
