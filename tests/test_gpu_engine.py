"""GPU parity, part 1: the row-solve engine behind spasm_schur / spasm_kernel / spasm_rref /
spasm_sparse_triangular_solve / spasm_transpose, called through the C ABI on host structs and
compared BIT-EXACTLY with the CPU oracle on the same seeded inputs.  The factor (U, qinv, p) fed to
both libraries is produced once by the oracle: the structs are plain host memory with the same
layout, so the same pointer is handed to both."""
import ctypes as C

import numpy as np
import pytest

import checks
import synth

pytestmark = pytest.mark.gpu


def make_fact(pkg, api, A, L=False):
    """empty spasm_lu the way spasm_echelonize allocates it (oracle/spasm_oracle.c)"""
    n, m = A.shape
    U = api.csr_alloc(n, m, max(A.nnz(), 1), A.prime)
    U.data.contents.n = 0
    qinv = np.full(m, -1, dtype=np.int32)
    lu = pkg._LU()
    lu.r, lu.complete = 0, 0
    lu.U = U.data
    lu.qinv = qinv.ctypes.data_as(C.POINTER(C.c_int32))
    lu.L, lu.p, lu.Ltmp = None, None, None
    return lu, U, qinv


def structural_round(pkg, oracle, A):
    lu, U, qinv = make_fact(pkg, oracle, A)
    p = np.zeros(max(A.n, 1), dtype=np.int32)
    opts = oracle.EchelonizeOpts()
    npiv = oracle.lib.spasm_pivots_extract_structural(A.data, None, C.byref(lu), p.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(opts))
    return lu, U, qinv, p, npiv


CASES = [
    # (n, m, k, prime, seed)   small primes exercise cancellations, the last one the 64-bit path
    (300, 320, 3, 42013, 1),
    (2000, 2000, 5, 42013, 2),   # rows fill in: medium / heavy tiers
    (1500, 1800, 4, 7, 3),
    (1200, 1000, 6, 65521, 4),
    (900, 900, 4, 4294967291, 5),
    (800, 850, 4, 2147483647, 6),
]


@pytest.mark.parametrize("case", CASES)
def test_transpose(gpu, oracle, case):
    n, m, k, prime, seed = case
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    Ao, Ag = oracle.from_arrays(n, m, p, j, x, prime), gpu.from_arrays(n, m, p, j, x, prime)
    To, Tg = oracle.transpose(Ao), gpu.transpose(Ag)
    for a, b in zip(To.arrays(), Tg.arrays()):
        assert np.array_equal(a, b)
    assert To.shape == Tg.shape == (m, n)


@pytest.mark.parametrize("case", CASES)
def test_schur_and_density(pkg, gpu, oracle, case):
    n, m, k, prime, seed = case
    p_, j_, x_ = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p_, j_, x_, prime)
    lu, U, qinv, p, npiv = structural_round(pkg, oracle, A)
    rest = np.ascontiguousarray(p[npiv:])
    nrest = n - npiv
    pp = rest.ctypes.data_as(C.POINTER(C.c_int32))
    outs = []
    for api in (oracle, gpu):
        pout = np.zeros(max(nrest, 1), dtype=np.int32)
        S = pkg.CSR(api, api.lib.spasm_schur(A.data, pp, nrest, C.byref(lu), 0.0, None, None, pout.ctypes.data_as(C.POINTER(C.c_int32))))
        outs.append((S.arrays(), pout[:nrest].copy(), S.shape))
    (so, po, sho), (sg, pg, shg) = outs
    # the algorithmic work behind the Schur GB/s figure: the kernels' own count (whichever stage solved a
    # row: shared-memory tier, global tier or the SpTRSM engine) equals the oracle's, byte for byte
    Ls = (C.c_longlong * 7)()
    gpu.lib.spasm_b200_last_stats.argtypes = [C.POINTER(C.c_longlong)]
    gpu.lib.spasm_b200_last_stats(Ls)
    ob = C.c_int64.in_dll(oracle.lib, "spasm_b200_last_bytes").value
    om = C.c_int64.in_dll(oracle.lib, "spasm_b200_last_macs").value
    assert (Ls[0], Ls[1]) == (ob, om), f"work counters: gpu {Ls[0]} B / {Ls[1]} MACs, oracle {ob} B / {om} MACs"
    assert sho == shg == (nrest, m)
    assert np.array_equal(po, pg)
    for a, b, name in zip(so, sg, "pjx"):
        assert np.array_equal(a, b), f"S.{name} differs"
    # nothing lands on a pivotal column
    assert (qinv[so[1]] < 0).all()
    # density estimate: same seeded sample, same count
    vals = []
    for api in (oracle, gpu):
        api.lib.spasm_b200_seed.argtypes = [C.c_uint64]
        api.lib.spasm_b200_seed(1234)
        vals.append(api.lib.spasm_schur_estimate_density(A.data, pp, nrest, U.data, qinv.ctypes.data_as(C.POINTER(C.c_int32)), 100))
    assert vals[0] == vals[1]


@pytest.mark.parametrize("case", CASES[:4])
def test_schur_with_L(pkg, gpu, oracle, case):
    n, m, k, prime, seed = case
    p_, j_, x_ = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p_, j_, x_, prime)
    lu, U, qinv, p, npiv = structural_round(pkg, oracle, A)
    rest = np.ascontiguousarray(p[npiv:])
    nrest = n - npiv
    res = []
    for api in (oracle, gpu):
        T = api.lib.spasm_triplet_alloc(n, n, 16, prime, True)
        S = pkg.CSR(api, api.lib.spasm_schur(A.data, rest.ctypes.data_as(C.POINTER(C.c_int32)), nrest, C.byref(lu), 0.0,
                                             C.cast(T, C.c_void_p), None, None))
        nz = T.contents.nz
        trip = (np.ctypeslib.as_array(T.contents.i, shape=(max(nz, 1),))[:nz].copy(),
                np.ctypeslib.as_array(T.contents.j, shape=(max(nz, 1),))[:nz].copy(),
                np.ctypeslib.as_array(T.contents.x, shape=(max(nz, 1),))[:nz].copy())
        api.lib.spasm_triplet_free(T)
        res.append((S.arrays(), trip))
    for a, b in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case", CASES)
def test_kernel_and_rref(pkg, gpu, oracle, case):
    n, m, k, prime, seed = case
    p_, j_, x_ = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p_, j_, x_, prime)
    fact = oracle.echelonize(A)
    Ko = oracle.kernel(fact)
    Kg = pkg.CSR(gpu, gpu.lib.spasm_kernel(fact.data))
    assert Ko.shape == Kg.shape
    for a, b, name in zip(Ko.arrays(), Kg.arrays(), "pjx"):
        assert np.array_equal(a, b), f"K.{name} differs"
    if n <= 1500:
        checks.check_kernel(gpu, A, fact, Kg)
    rq_o, rq_g = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
    Ro = oracle.rref(fact, rq_o)
    Rg = pkg.CSR(gpu, gpu.lib.spasm_rref(fact.data, rq_g.ctypes.data_as(C.POINTER(C.c_int32))))
    assert np.array_equal(rq_o, rq_g)
    for a, b, name in zip(Ro.arrays(), Rg.arrays(), "pjx"):
        assert np.array_equal(a, b), f"R.{name} differs"


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[4]])
def test_sparse_triangular_solve(pkg, gpu, oracle, case):
    """src/SpaSM.jl:694-713: same top, same pattern SET, same x on the pattern; x_b.U + x_a == B[k]"""
    n, m, k, prime, seed = case
    p_, j_, x_ = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p_, j_, x_, prime)
    fact = oracle.echelonize(A, max_round=1, enable_dense=False)
    U, qinv = fact.U, fact.qinv.copy()
    for row in range(0, n, max(1, n // 12)):
        res = []
        for api in (oracle, gpu):
            xj = np.zeros(3 * m, dtype=np.int32)
            x = np.full(m, 777, dtype=np.int32)
            top = api.sparse_triangular_solve_row(U, A, row, xj, x, qinv)
            pat = xj[top:m].copy()
            res.append((top, set(pat.tolist()), {int(c): int(x[c]) for c in pat}, pat))
        assert res[0][0] == res[1][0], "top differs"
        assert res[0][1] == res[1][1], "pattern set differs"
        assert res[0][2] == res[1][2], "x differs on the pattern"
        # the GPU pattern is a topological order: a pivotal column comes before the pivotal columns its row holds
        pat = res[1][3]
        posn = {int(c): t for t, c in enumerate(pat)}
        Up, Uj, _ = U.arrays()
        for c in pat:
            i = qinv[c]
            if i >= 0:
                for cc in Uj[Up[i] : Up[i + 1]]:
                    if cc != c and qinv[cc] >= 0:
                        assert posn[int(cc)] > posn[int(c)]


def _triplets(api, n, m, nz, prime, seed):
    """a triplet list built directly in the library's host arrays: many duplicate (row, column) pairs, sums that vanish
    (small primes), empty rows and rows far longer than 64 entries"""
    rng = np.random.default_rng(seed)
    T = api.lib.spasm_triplet_alloc(n, m, max(nz, 1), prime, True)
    if nz:
        ti = rng.integers(0, n, size=nz)
        ti[rng.random(nz) < 0.3] = rng.integers(0, max(1, n // 50), size=int((rng.random(nz) < 0.3).sum()) or 1)[0]  # one crowded row
        tj = rng.integers(0, m, size=nz)
        dup = rng.random(nz) < 0.5
        tj[dup] = rng.integers(0, min(m, 37), size=int(dup.sum()))  # collisions
        tx = synth.balanced(rng.integers(1, prime, size=nz), prime)
        np.ctypeslib.as_array(T.contents.i, shape=(nz,))[:] = ti
        np.ctypeslib.as_array(T.contents.j, shape=(nz,))[:] = tj
        np.ctypeslib.as_array(T.contents.x, shape=(nz,))[:] = tx
    T.contents.nz = nz
    return T


@pytest.mark.parametrize("case", [(50, 40, 400, 7, 1), (2000, 3000, 30000, 42013, 2), (300, 5000, 60000, 3, 3), (1000, 800, 0, 42013, 4),
                                  (5, 100000, 3000, 65521, 5), (4000, 4000, 20000, 4294967291, 6)])
def test_compress_on_device(pkg, gpu, oracle, case):
    """triplets -> CSR on the device (csrc/compress.cu; spasm_compress, src/SpaSM.jl:479-493): the same arrays, bit for bit, as the
    oracle's and the library's own host code — rows in the order of the triplet list, duplicates summed into the first
    occurrence, zero sums dropped"""
    n, m, nz, prime, seed = case
    T = _triplets(oracle, n, m, nz, prime, seed)
    ref, host, dev = oracle.compress(T), gpu.compress(T), gpu.compress(T, device=True)
    assert ref.shape == host.shape == dev.shape == (n, m)
    for a, b, c, name in zip(ref.arrays(), host.arrays(), dev.arrays(), "pjx"):
        assert np.array_equal(a, b), f"host compress: {name} differs from the oracle"
        assert np.array_equal(a, c), f"device compress: {name} differs from the oracle"
    if nz:
        assert ref.nnz() < nz  # duplicates / cancellations really occurred
    oracle.lib.spasm_triplet_free(T)
