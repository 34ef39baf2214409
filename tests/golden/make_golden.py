"""Generates the committed golden fixtures from the reference's own test data.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Outputs (small, committed):
  adeno_fiber_lcs.npz   sequences of test/adeno_fiber/adeno_fiber + the exact LCS lengths pinned by
                        test/adeno_fiber/pid_sq.csv (pid = lcs / min(len), 6 decimals; row = seq0)
                        + the distances of dist_sq.csv (indel075_div_lcs, float, 6 decimals)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from famsa_b200 import seqio  # noqa: E402

REF = "/root/reference/test"


def read_square(path):
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split(",")[1:]
        rows = []
        for line in fh:
            parts = line.rstrip("\n").split(",")
            rows.append([float(x) for x in parts[1:]])
    return header, np.array(rows, dtype=np.float64)


def main():
    ids, seqs = seqio.read_fasta(os.path.join(REF, "adeno_fiber", "adeno_fiber"))
    codes = [seqio.encode(s) for s in seqs]
    lens = np.array([len(c) for c in codes])
    hdr, pid = read_square(os.path.join(REF, "adeno_fiber", "pid_sq.csv"))
    assert hdr == ids and pid.shape == (len(ids), len(ids))
    minlen = np.minimum(lens[:, None], lens[None, :])
    lcs = np.rint(pid * minlen).astype(np.uint16)
    # the 6-decimal print must round-trip, otherwise the fixture would not pin the integer
    assert np.all(np.abs(lcs / minlen - pid) < 6e-7)
    _, dist = read_square(os.path.join(REF, "adeno_fiber", "dist_sq.csv"))
    np.savez_compressed(os.path.join(HERE, "adeno_fiber_lcs.npz"),
                        seqs=np.array(seqs), ids=np.array(ids), lcs=lcs, dist=dist.astype(np.float64))
    print("adeno_fiber_lcs.npz:", lcs.shape, "lcs range", lcs.min(), lcs.max())


if __name__ == "__main__":
    main()
