"""GPU parity, part 2: the whole path through the C ABI — spasm_echelonize (all option paths),
spasm_pivots_extract_structural, spasm_dense_rref, spasm_kernel, spasm_solve / spasm_gesv, SpMV —
bit-exact against the CPU oracle on the same seeded inputs, plus the canonical invariants computed
independently of the oracle."""
import ctypes as C

import numpy as np
import pytest

import checks
import synth
from test_gpu_engine import make_fact

pytestmark = pytest.mark.gpu

CASES = [
    # n, m, k, prime, seed, ragged
    (60, 60, 3, 42013, 1, False),
    (400, 380, 3, 42013, 2, False),
    (1500, 1500, 5, 42013, 3, False),   # goes dense after round 0 (like C1)
    (1200, 2400, 4, 42013, 4, False),   # wide (like C5)
    (2500, 900, 4, 65521, 5, False),    # tall
    (700, 800, 6, 7, 6, True),          # tiny prime: cancellations; ragged incl. empty rows
    (600, 600, 4, 3, 7, True),
    (500, 520, 4, 4294967291, 8, False),  # largest prime allowed (src/SpaSM.jl:74): 64-bit path
    (500, 480, 4, 2147483647, 9, True),
]


def _mk(api, case):
    n, m, k, prime, seed, ragged = case
    p, j, x = (synth.ragged_rows if ragged else synth.random_rows)(n, m, k, prime, seed)
    return api.from_arrays(n, m, p, j, x, prime)


@pytest.mark.parametrize("case", CASES)
def test_structural_pivots(pkg, gpu, oracle, case):
    """FL + FL-on-columns + greedy search + reorder + extraction == sequential oracle"""
    A = _mk(oracle, case)
    res = []
    for api in (oracle, gpu):
        lu, U, qinv = make_fact(pkg, api, A)
        p = np.zeros(max(A.n, 1), dtype=np.int32)
        opts = api.EchelonizeOpts()
        npiv = api.lib.spasm_pivots_extract_structural(A.data, None, C.byref(lu), p.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(opts))
        res.append(dict(npiv=npiv, p=p.copy(), qinv=qinv.copy(), Up=U.arrays()[0], Uj=U.arrays()[1], Ux=U.arrays()[2]))
    checks.assert_same(res[0], res[1], "pivots: ")


OPTS = [
    dict(),
    dict(enable_dense=False),                      # sparse rounds + GPLU
    dict(enable_dense=False, enable_greedy_pivot_search=False),
    dict(max_round=0, enable_dense=False),         # pure GPLU
    dict(sparsity_threshold=0.0, max_round=1, dense_block_size=64),  # forced dense, several blocks
    dict(max_round=1),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kw", OPTS)
def test_echelonize_bit_exact(gpu, oracle, case, kw):
    A = _mk(oracle, case)
    fo = oracle.echelonize(A, **kw)
    fg = gpu.echelonize(A, **kw)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg), f"{kw}: ")


@pytest.mark.parametrize("case", CASES[:6])
def test_invariants_and_kernel(gpu, oracle, case):
    A = _mk(gpu, case)
    fact = gpu.echelonize(A)
    checks.check_U_structure(gpu, fact)
    checks.check_rank_and_rowspace(gpu, A, fact)
    K = gpu.kernel(fact)
    checks.check_kernel(gpu, A, fact, K)
    Ko = oracle.kernel(oracle.echelonize(A))
    for a, b in zip(K.arrays(), Ko.arrays()):
        assert np.array_equal(a, b)
    Rq = np.zeros(A.m, dtype=np.int32)
    R = gpu.rref(fact, Rq)
    checks.check_rref(gpu, fact, R, Rq)


def test_reference_goldens_on_gpu(gpu):
    """test/runtests.jl:7-24 and README.md:9-48 through the CUDA library"""
    import json
    from pathlib import Path

    import scipy.sparse as sp

    G = json.loads((Path(__file__).parent / "golden" / "reference_goldens.json").read_text())
    g = G["runtests"]
    p = g["prime"]
    m = sp.csc_matrix((g["V"], (np.array(g["I"]) - 1, np.array(g["J"]) - 1)))
    sm = gpu.CSR(m)
    assert (gpu.sparse(gpu.transpose(gpu.transpose(sm))) != gpu.sparse(sm)).nnz == 0
    for key, mat in (("kernel", sm), ("kernel_transpose", gpu.transpose(sm))):
        kk = g[key]
        want = sp.csc_matrix((kk["V"], (np.array(kk["I"]) - 1, np.array(kk["J"]) - 1)), shape=tuple(kk["shape"]))
        got = gpu.sparse(gpu.kernel(mat))
        assert got.shape == want.shape and not ((got - want).toarray() % p).any()
    r = G["readme"]
    sm = gpu.CSR(sp.csc_matrix((r["V"], (np.array(r["I"]) - 1, np.array(r["J"]) - 1))))
    lines = []
    gpu.log(lambda s: lines.append(s) or 0)
    try:
        fact = gpu.echelonize(sm, verbose=True)
        k = gpu.kernel(fact, verbose=True)
    finally:
        gpu.log(None)
    assert (fact.r, fact.U.nnz(), k.nnz()) == (r["rank"], r["nz_in_basis"], r["nnz_K"])
    text = "".join(lines)
    for must in r["log_must_contain"]:
        assert must in text, must


@pytest.mark.parametrize("prime", [7, 42013, 65521, 2147483647, 4294967291])
def test_dense_rref(gpu, oracle, prime):
    for (n, m, rank, seed) in [(50, 80, None, 1), (200, 150, 70, 2), (130, 130, None, 3), (1, 9, None, 4)]:
        D = synth.dense_random(n, m, prime, seed, rank)
        a, b = D.copy(), D.copy()
        ro, po = oracle.dense_rref(prime, a)
        rg, pg = gpu.dense_rref(prime, b)
        assert ro == rg and np.array_equal(po, pg)
        assert np.array_equal(a[:ro], b[:rg])
        if rank is not None:
            assert rg == min(rank, n, m)


@pytest.mark.parametrize("prime", [251, 42013, 4294967291])
def test_solve_gesv_spmv(gpu, oracle, prime):
    n, m, k = 300, 420, 4
    p, j, x = synth.random_rows(n, m, k, prime, 21)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    fo, fg = oracle.echelonize(A, L=True), gpu.echelonize(A, L=True)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg), "L=true: ")
    Ad = checks.dense_of(gpu, A)
    assert np.array_equal(checks.mm(checks.dense_of(gpu, fg.L), checks.dense_of(gpu, fg.U), prime), Ad), "A != L.U"
    rng = np.random.default_rng(3)
    for t in range(5):
        x0 = rng.integers(0, prime, size=n)
        b = checks.mm(x0, Ad, prime) if t < 3 else rng.integers(0, prime, size=m)
        bb = synth.balanced(b, prime)
        xo, xg = oracle.solve(fo, bb), gpu.solve(fg, bb)
        assert (xo is None) == (xg is None)
        if xg is not None:
            assert np.array_equal(xo, xg)
            assert np.array_equal(checks.mm(xg.astype(np.int64) % prime, Ad, prime), b % prime)
    # SpMV
    xv = synth.balanced(rng.integers(0, prime, size=n), prime)
    yv = synth.balanced(rng.integers(0, prime, size=m), prime)
    assert np.array_equal(gpu.xapy(xv, A, yv.copy()), oracle.xapy(xv, A, yv.copy()))
    assert np.array_equal(gpu.axpy(A, yv, xv.copy()), oracle.axpy(A, yv, xv.copy()))
    # gesv
    import scipy.sparse as sp

    Bd = np.vstack([Ad[3], (Ad[1] + 2 * Ad[2]) % prime, rng.integers(0, prime, size=m)])
    B = gpu.CSR(sp.csc_matrix(Bd.T), prime)
    Xo, oko = oracle.gesv(fo, B)
    Xg, okg = gpu.gesv(fg, B)
    assert np.array_equal(oko, okg)
    for a, b2 in zip(Xo.arrays(), Xg.arrays()):
        assert np.array_equal(a, b2)


def test_gesv_many_right_hand_sides(gpu, oracle):
    """the batched gesv (all right-hand sides through the row engine at once): 40 systems, solvable ones, random ones,
    a zero row and a repeated row, against the oracle's one-at-a-time loop — ok[] and X bit for bit, and x.A == b"""
    import scipy.sparse as sp

    prime, n, m = 42013, 500, 800
    p, j, x = synth.random_rows(n, m, 5, prime, 33)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    fo, fg = oracle.echelonize(A, L=True), gpu.echelonize(A, L=True)
    Ad = checks.dense_of(gpu, A)
    rng = np.random.default_rng(8)
    rows = []
    for t in range(40):
        if t % 5 == 4:
            rows.append(rng.integers(0, prime, size=m))  # almost surely not in the row space
        elif t == 7:
            rows.append(np.zeros(m, dtype=np.int64))
        else:
            x0 = np.zeros(n, dtype=np.int64)
            idx = rng.choice(n, size=1 + t % 17, replace=False)
            x0[idx] = rng.integers(1, prime, size=len(idx))
            rows.append(checks.mm(x0, Ad, prime))
    rows[11] = rows[10].copy()
    Bd = np.vstack(rows)
    B = gpu.CSR(sp.csc_matrix(Bd.T), prime)
    Xo, oko = oracle.gesv(fo, B)
    Xg, okg = gpu.gesv(fg, B)
    assert np.array_equal(oko, okg) and okg.sum() >= 30
    for a, b2 in zip(Xo.arrays(), Xg.arrays()):
        assert np.array_equal(a, b2)
    Xd = checks.dense_of(gpu, Xg)
    good = np.nonzero(okg)[0]
    assert np.array_equal(checks.mm(Xd[good], Ad, prime), Bd[good] % prime)


def test_medium_scale_c1_like(gpu, oracle):
    """a down-scaled configs[0] (random 5 nnz/row, mod 42013): echelonize + kernel, bit-exact"""
    n = m = 3000
    p, j, x = synth.random_rows(n, m, 5, 42013, 0x5A5A0001)
    A = gpu.from_arrays(n, m, p, j, x, 42013)
    fo, fg = oracle.echelonize(A), gpu.echelonize(A)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg))
    Ko, Kg = oracle.kernel(fo), gpu.kernel(fg)
    for a, b in zip(Ko.arrays(), Kg.arrays()):
        assert np.array_equal(a, b)


def test_gl7d19_shaped_sparse_regime(gpu, oracle):
    """configs[2] at 1/128 scale (banded planted-rank generator): stays sparse, so the rounds of
    structural pivots + sparse Schur complement + GPLU tail all run; rank known by construction;
    every array bit-exact against the oracle"""
    s = 128
    n, m, r = 1911130 // s, 1955309 // s, 1033568 // s
    p, j, x = synth.banded_planted(n, m, r, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
    A = gpu.from_arrays(n, m, p, j, x, 42013)
    fo, fg = oracle.echelonize(A), gpu.echelonize(A)
    assert fg.r == r
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg))
    checks.check_U_structure(gpu, fg)
    fo2, fg2 = oracle.echelonize(A, L=True), gpu.echelonize(A, L=True)
    checks.assert_same(checks.lu_arrays(fo2), checks.lu_arrays(fg2), "L: ")


def test_dense_tail_entry_point(gpu):
    """BASELINE configs[3] entry point at small size: a random dense n x n matrix mod 65521 has full rank
    (probability 1 - 1/p) and the blocked elimination must find it, for several block sizes"""
    f = gpu.lib.spasm_b200_dense_tail_bench
    f.restype = C.c_int
    f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
    ms = C.c_double(0)
    for (n, m, block) in [(700, 700, 256), (1500, 1200, 1000), (900, 2000, 300)]:
        assert f(65521, n, m, block, 1234, C.byref(ms)) == min(n, m)


def test_multi_gpu_sharded_tail():
    """torchrun with 2 ranks when the box has >= 2 GPUs (the driver's single-GPU run skips this)"""
    import subprocess
    import sys
    from pathlib import Path

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", "tests/dist_gpu_check.py"], cwd=str(root), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dist check OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


LOWRANK = [
    # n, m, k, prime, seed, options
    (3000, 200, 3, 42013, 5, {}),
    (2500, 150, 2, 7, 6, dict(dense_block_size=64)),
    (4000, 300, 3, 65521, 7, dict(low_rank_start_weight=4, dense_block_size=100)),
    (800, 120, 2, 4294967291, 8, dict(dense_block_size=50)),
    (1500, 90, 3, 42013, 9, dict(max_round=0, dense_block_size=16)),                              # every pivot through the low-rank blocks
    (1500, 90, 3, 251, 10, dict(max_round=0, dense_block_size=16, low_rank_start_weight=1)),      # weight doubling
    (20000, 60, 2, 42013, 11, dict(max_round=0, dense_block_size=32)),                            # K > 16384: chunked tensor-core GEMM
]


@pytest.mark.parametrize("case", LOWRANK)
def test_low_rank_mode(gpu, oracle, case):
    """tall-and-skinny finish (spasm_schur_dense_randomized, src/SpaSM.jl:767-769): random combinations of the
    remaining rows, same counter-based random numbers on both sides -> bit-exact"""
    n, m, k, prime, seed, kw = case
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    lines = []
    oracle.log(lambda s: lines.append(s) or 0)
    try:
        fo = oracle.echelonize(A, verbose=True, **kw)
    finally:
        oracle.log(None)
    assert any("low-rank" in l for l in lines)
    fg = gpu.echelonize(A, **kw)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg), f"{kw}: ")
    checks.check_U_structure(gpu, fg)
    if n <= 4000:
        checks.check_rank_and_rowspace(gpu, A, fg)


@pytest.mark.gpu
@pytest.mark.parametrize("case", __import__("test_oracle_invariants").MIDTAIL + [
    (6000, 6000, 3, 42013, 21, 2300, dict(sparsity_threshold=0.0, max_round=0, dense_block_size=500)),  # deferred far rows pending at the switch
])
def test_mid_tail_low_rank_switch(gpu, oracle, case):
    """the dense loop meets a block with fewer than low_rank_ratio * Sn pivots and hands the remaining
    rows to the low-rank mode (random combinations of the rows of the dense Schur complement it already holds)"""
    import ctypes as C
    from test_oracle_invariants import midtail_input

    A, kw = midtail_input(gpu, case)
    fo = oracle.echelonize(A, **kw)
    gpu.lib.spasm_b200_lowrank_switches.restype = C.c_longlong
    gpu.lib.spasm_b200_lowrank_switches.argtypes = [C.c_int]
    gpu.lib.spasm_b200_lowrank_switches(1)
    fg = gpu.echelonize(A, **kw)
    assert gpu.lib.spasm_b200_lowrank_switches(1) == 1
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg), f"{kw}: ")
    checks.check_U_structure(gpu, fg)


@pytest.mark.gpu
@pytest.mark.parametrize("case", __import__("test_oracle_invariants").DENSE_L + [
    (3000, 3000, 8, 42013, 31, dict(dense_block_size=300)),      # several panels, tcgen05 products
    (2600, 2400, 6, 2147483647, 32, dict(dense_block_size=256)),
    (5200, 5000, 9, 65521, 33, dict(dense_block_size=500)),       # deferred far rows (default depth) under the with-L mode
])
def test_dense_tail_with_L(gpu, oracle, case):
    """echelonize(L=true) on a matrix that goes dense: the tail runs on the tensor cores (row echelon form recovered from the
    reduced panels, multipliers through the same limb GEMM) and equals the oracle's restatement of spasm_ffpack_LU
    (src/SpaSM.jl:806) bit for bit: U, L, Lp, qinv; A = L.U; solve / gesv work on the factor"""
    from test_oracle_invariants import check_A_equals_LU

    n, m, k, prime, seed, kw = case
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    gpu.lib.spasm_b200_mma_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    st = (C.c_double * 4)()
    gpu.lib.spasm_b200_mma_stats(st, 1)
    fg = gpu.echelonize(A, L=True, **kw)
    gpu.lib.spasm_b200_mma_stats(st, 0)
    if n >= 2600:
        assert st[2] > 0, "the tensor-core kernel did not run for the with-L tail"
    fo = oracle.echelonize(A, L=True, **kw)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg), f"L, {kw}: ")
    if n <= 900:
        check_A_equals_LU(gpu, (n, m, p, j, x, prime), fg)
    rng = np.random.default_rng(seed)
    xv = np.zeros(n, dtype=np.int64)
    xv[rng.integers(0, n, size=20)] = rng.integers(1, prime, size=20)
    Ad_rows = {}
    b = np.zeros(m, dtype=object)
    for i in np.nonzero(xv)[0]:
        for e in range(p[i], p[i + 1]):
            b[j[e]] = (b[j[e]] + int(xv[i]) * int(x[e])) % prime
    bb = synth.balanced(np.array(b, dtype=np.int64), prime)
    sg, so = gpu.solve(fg, bb), oracle.solve(fo, bb)
    assert sg is not None and so is not None and np.array_equal(np.asarray(sg), np.asarray(so))


def test_blocks_gpu(gpu):
    from test_blocks import check_blocks

    check_blocks(gpu)
