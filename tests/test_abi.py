"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/famsa_b200.h declares, fails loudly (no CPU fallback) when there is no device, and the
host-side Transform matches the oracle."""
import os
import re

import numpy as np
import pytest

import famsa_b200
from conftest import ROOT
from oracle import pyoracle


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "famsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(famsa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = famsa_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/famsa_b200.h but not exported"
    assert sorted(famsa_b200.EXPORTED_SYMBOLS) == syms
    assert lib.famsa_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(famsa_b200.FamsaError, match="no CPU fallback"):
        famsa_b200.Engine(0)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (see tier rules)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "famsa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/ (", ""), f"{f} mentions the oracle"


def test_host_transform_matches_oracle():
    lib = famsa_b200.load_library()
    rng = np.random.default_rng(2)
    for _ in range(300):
        l1, l2 = (int(x) for x in rng.integers(1, 700, size=2))
        lcs = int(rng.integers(0, min(l1, l2) + 1))
        for kind in (0, 1, 2):
            assert lib.famsa_transform_f64(kind, lcs, l1, l2) == pyoracle.transform(kind, lcs, l1, l2, True)
            assert lib.famsa_transform_f32(kind, lcs, l1, l2) == pyoracle.transform(kind, lcs, l1, l2, False)
