"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle would need
hours here): rank(A) = rank(A^T) through two unrelated pivot sequences, the rank of a matrix does not change
when its rows are permuted, and the dense tail finds a rank that is known by construction."""
import ctypes as C

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def resident_rank(gpu, A, **kw):
    """rank through the device-resident entry points bench.py times (the 48 GB factor is not downloaded)"""
    lib = gpu.lib
    lib.spasm_b200_upload.restype = C.c_void_p
    lib.spasm_b200_upload.argtypes = [C.c_void_p]
    lib.spasm_b200_echelonize_resident.restype = C.c_int
    lib.spasm_b200_echelonize_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.spasm_b200_release.argtypes = [C.c_void_p]
    h = lib.spasm_b200_upload(C.cast(A.data, C.c_void_p))
    assert h
    try:
        opts = gpu.EchelonizeOpts(**kw)
        ms = C.c_double(0)
        r = lib.spasm_b200_echelonize_resident(h, C.byref(opts), C.byref(ms))
        assert r >= 0
        return r
    finally:
        lib.spasm_b200_release(h)


def test_configs1_full_size_rank_properties(gpu):
    """configs[1]: 200 000 x 200 000, 10 non-zeros per row, mod 42013 (the matrix bench.py times)"""
    import bench

    n = bench.FULL_N
    p, j, x = bench.make_input(n)
    A = gpu.from_arrays(n, n, p, j, x, bench.PRIME)
    r = resident_rank(gpu, A)
    # a uniform random sparse matrix of this density is singular only through its empty / duplicated columns
    assert n - 64 < r <= n
    At = gpu.transpose(A)
    assert resident_rank(gpu, At) == r
    # rows in another order: other pivots, other Schur complement, same rank
    perm = np.random.default_rng(7).permutation(n)
    lens = np.diff(p)[perm]
    pp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = (np.repeat(p[:-1][perm] - pp[:-1], lens) + np.arange(pp[-1])).astype(np.int64)
    Ap = gpu.from_arrays(n, n, pp, j[idx], x[idx], bench.PRIME)
    assert resident_rank(gpu, Ap) == r


def test_configs3_full_size_dense_tail(gpu):
    """configs[3]: dense 32768 x 32768 mod 65521 into the dense-tail entry point; full rank for the iid matrix
    (probability 1 - 1.5e-5), planted rank 24576 for the product of two random factors"""
    lib = gpu.lib
    f = lib.spasm_b200_dense_tail_bench
    f.restype = C.c_int
    f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
    g = lib.spasm_b200_dense_tail_bench_planted
    g.restype = C.c_int
    g.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
    ms = C.c_double(0)
    n = 32768
    assert f(65521, n, n, 1000, 0x5A5A0004, C.byref(ms)) == n
    assert g(65521, n, n, 24576, 1000, 0x5A5A0004, C.byref(ms)) == 24576
    # small planted cases, several block sizes, rank below / above one block
    for (a, b, r, block) in [(900, 700, 130, 256), (1500, 1500, 999, 1000), (2000, 1200, 1100, 300)]:
        assert g(65521, a, b, r, block, 99, C.byref(ms)) == r
