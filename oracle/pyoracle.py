"""ctypes access to the checkers: oracle/liboracle.so (our C restatement) and
oracle/_ref/libfamsa_ref.so (the unmodified reference behind oracle/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from famsa_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libfamsa_ref.so")


def build(quiet: bool = True):
    """(Re)build the checkers: liboracle.so always, _ref only where /root/reference exists."""
    subprocess.run(["make", "-C", _HERE, "-j8", "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_oracle = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        vp, u32 = C.c_void_p, C.c_uint32
        lib.lcs_oracle_rows.argtypes = [vp, vp, vp, vp, u32, vp, u32, vp]
        lib.lcs_oracle_triangle.argtypes = [vp, vp, vp, u32, u32, vp]
        lib.lcs_oracle_transform_f64.argtypes = [C.c_int, u32, u32, u32]
        lib.lcs_oracle_transform_f64.restype = C.c_double
        lib.lcs_oracle_transform_f32.argtypes = [C.c_int, u32, u32, u32]
        lib.lcs_oracle_transform_f32.restype = C.c_float
        _oracle = lib
    return _oracle


def lcs_rows(codes, offsets, lens, ref_ids, col_ids=None, n_col=None) -> np.ndarray:
    lib = oracle()
    ref = np.ascontiguousarray(ref_ids, dtype=np.uint32)
    cols = None if col_ids is None else np.ascontiguousarray(col_ids, dtype=np.uint32)
    n_col = (len(lens) if n_col is None else n_col) if cols is None else len(cols)
    out = np.zeros(max(len(ref) * n_col, 1), dtype=np.uint32)
    lib.lcs_oracle_rows(_p(codes), _p(offsets), _p(lens), _p(ref), len(ref), _p(cols), n_col, _p(out))
    return out[:len(ref) * n_col].reshape(len(ref), n_col)


def lcs_triangle(codes, offsets, lens, row_begin=0, row_end=None) -> np.ndarray:
    lib = oracle()
    row_end = len(lens) if row_end is None else row_end
    f = lambda r: r * (r - 1) // 2 if r else 0
    size = f(row_end) - f(row_begin)
    out = np.zeros(max(size, 1), dtype=np.uint32)
    lib.lcs_oracle_triangle(_p(codes), _p(offsets), _p(lens), row_begin, row_end, _p(out))
    return out[:size]


def transform(kind: int, lcs: int, len1: int, len2: int, double=True) -> float:
    lib = oracle()
    return float((lib.lcs_oracle_transform_f64 if double else lib.lcs_oracle_transform_f32)(kind, lcs, len1, len2))


# ------------------------------------------------------------------ the real reference
_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        vp, u32 = C.c_void_p, C.c_uint32
        lib.ref_seqset_create.argtypes = [vp, u32]
        lib.ref_seqset_create.restype = vp
        lib.ref_seqset_destroy.argtypes = [vp]
        lib.ref_seqset_len.argtypes = [vp, u32]
        lib.ref_seqset_len.restype = u32
        lib.ref_seqset_codes.argtypes = [vp, u32, vp]
        lib.ref_lcs_row_prefix.argtypes = [vp, u32, u32, vp, C.c_int]
        lib.ref_lcs_row_ids.argtypes = [vp, u32, vp, u32, vp, C.c_int]
        lib.ref_lcs_triangle_mt.argtypes = [vp, u32, u32, C.c_int, C.c_int, vp, vp]
        lib.ref_lcs_triangle_mt.restype = C.c_double
        lib.ref_transform_f64.argtypes = [C.c_int, u32, u32, u32]
        lib.ref_transform_f64.restype = C.c_double
        lib.ref_transform_f32.argtypes = [C.c_int, u32, u32, u32]
        lib.ref_transform_f32.restype = C.c_float
        _ref = lib
    return _ref


class RefSeqSet:
    """A set of CSequence objects inside the reference, built from residue letters."""

    def __init__(self, letters: list[str]):
        self.lib = ref()
        arr = (C.c_char_p * len(letters))(*[s.encode("ascii") for s in letters])
        self.n = len(letters)
        self.h = C.c_void_p(self.lib.ref_seqset_create(arr, self.n))

    def close(self):
        if self.h:
            self.lib.ref_seqset_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def codes(self) -> list[np.ndarray]:
        out = []
        for i in range(self.n):
            ln = self.lib.ref_seqset_len(self.h, i)
            a = np.zeros(max(ln, 1), dtype=np.int8)
            self.lib.ref_seqset_codes(self.h, i, _p(a))
            out.append(a[:ln])
        return out

    def row_prefix(self, ref_id: int, n_cols: int, isa: int = 2) -> np.ndarray:
        out = np.zeros(max(n_cols, 1), dtype=np.uint32)
        self.lib.ref_lcs_row_prefix(self.h, ref_id, n_cols, _p(out), isa)
        return out[:n_cols]

    def row_ids(self, ref_id: int, ids, isa: int = 2) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.zeros(max(len(ids), 1), dtype=np.uint32)
        self.lib.ref_lcs_row_ids(self.h, ref_id, _p(ids), len(ids), _p(out), isa)
        return out[:len(ids)]

    def triangle_mt(self, row_begin: int, row_end: int, n_threads: int, isa: int = 2, want_lcs: bool = False):
        f = lambda r: r * (r - 1) // 2 if r else 0
        out = np.zeros(max(f(row_end) - f(row_begin), 1), dtype=np.uint32) if want_lcs else None
        pairs = C.c_uint64()
        sec = self.lib.ref_lcs_triangle_mt(self.h, row_begin, row_end, n_threads, isa, _p(out), C.byref(pairs))
        return sec, pairs.value, (out[:f(row_end) - f(row_begin)] if want_lcs else None)
