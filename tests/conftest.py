import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def adeno():
    """test/adeno_fiber of the reference: sequences + golden exact LCS matrix (pid_sq.csv)."""
    from famsa_b200 import seqio
    z = np.load(os.path.join(GOLDEN, "adeno_fiber_lcs.npz"))
    seqs = [str(s) for s in z["seqs"]]
    code_list = [seqio.encode(s) for s in seqs]
    codes, offsets, lens = seqio.pack(code_list)
    return dict(seqs=seqs, code_list=code_list, codes=codes, offsets=offsets, lens=lens,
                lcs=z["lcs"].astype(np.uint32), dist=z["dist"])


@pytest.fixture(scope="session")
def engine():
    import famsa_b200
    eng = famsa_b200.Engine(0)
    yield eng
    eng.close()


QUIRK_SEQS = ["A", "A" * 192, "AC", "A" * 128]
# pid_sq of the reference on QUIRK_SEQS (row = seq0) -> LCS lengths
QUIRK_LCS = np.array([[1, 1, 1, 1],
                      [2, 192, 2, 128],
                      [1, 1, 2, 1],
                      [1, 128, 1, 128]], dtype=np.uint32)


def random_set(rng, n, lo, hi, alphabet=20, with_specials=True):
    from famsa_b200 import seqio
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi + 1))
        c = rng.integers(0, alphabet, size=ln).astype(np.int8)
        if with_specials and ln:
            k = rng.random(ln) < 0.03
            c[k] = rng.integers(20, 24, size=int(k.sum())).astype(np.int8)
        out.append(c)
    return out
