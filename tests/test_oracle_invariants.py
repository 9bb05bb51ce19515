"""The oracle itself must satisfy the canonical invariants (SURVEY.md §8c level 3/4) on many
shapes, primes and option paths — this is what gives the unpinned restatement its credibility."""
import numpy as np
import pytest

import checks
import synth

PRIMES = [3, 7, 251, 42013, 65521, 2147483647, 4294967291]


def _mk(api, n, m, k, prime, seed, ragged=False):
    p, j, x = (synth.ragged_rows if ragged else synth.random_rows)(n, m, k, prime, seed)
    return api.from_arrays(n, m, p, j, x, prime)


@pytest.mark.parametrize("prime", PRIMES)
def test_field_ops(oracle, pkg, prime):
    import ctypes as C

    F = pkg._Field()
    oracle.lib.spasm_field_init(prime, C.byref(F))
    assert (F.p, F.halfp, F.mhalfp) == (prime, prime // 2, prime // 2 - prime + 1)
    rng = np.random.default_rng(prime)
    vals = [0, 1, -1, F.halfp, F.mhalfp] + rng.integers(F.mhalfp, F.halfp + 1, size=200).tolist()
    for a in vals[:40]:
        for b in vals[:40]:
            want = pkg.ZZp(prime, a * b).v
            assert oracle.lib.spasm_ZZp_mul(C.byref(F), a, b) == want
            assert oracle.lib.spasm_ZZp_axpy(C.byref(F), a, b, vals[7]) == pkg.ZZp(prime, a * b + vals[7]).v
            assert oracle.lib.spasm_ZZp_add(C.byref(F), a, b) == pkg.ZZp(prime, a + b).v
        if a % prime:
            inv = oracle.lib.spasm_ZZp_inverse(C.byref(F), a)
            assert (inv * a) % prime == 1 and F.mhalfp <= inv <= F.halfp


@pytest.mark.parametrize("prime", PRIMES)
@pytest.mark.parametrize("shape", [(40, 40, 3), (60, 35, 4), (30, 70, 5), (80, 80, 2)])
def test_rank_rowspace_kernel(oracle, prime, shape):
    n, m, k = shape
    for seed in range(3):
        A = _mk(oracle, n, m, k, prime, seed + 10 * n, ragged=(seed == 2))
        fact = oracle.echelonize(A)
        checks.check_U_structure(oracle, fact)
        checks.check_rank_and_rowspace(oracle, A, fact)
        K = oracle.kernel(fact)
        checks.check_kernel(oracle, A, fact, K)
        Rq = np.zeros(m, dtype=np.int32)
        R = oracle.rref(fact, Rq)
        checks.check_rref(oracle, fact, R, Rq)


def test_edge_cases(oracle):
    prime = 42013
    # empty matrix, zero rows, zero columns used, single entry
    for n, m in [(5, 7), (1, 1), (7, 5)]:
        Z = oracle.spzeros(oracle.CSR(np.zeros((1, 1))).field, n, m)
        fact = oracle.echelonize(Z)
        assert fact.r == 0
        K = oracle.kernel(fact)
        assert K.shape == (m, m) and K.nnz() == m
    A = oracle.CSR(np.array([[0, 0, 5], [0, 0, 0]]))  # SpaSM 3x2, one entry
    fact = oracle.echelonize(A)
    assert fact.r == 1
    # identity, full rank: empty kernel
    I = oracle.CSR(np.eye(6, dtype=np.int64))
    fact = oracle.echelonize(I)
    assert fact.r == 6 and oracle.kernel(fact).shape == (0, 6)
    # values that reduce to zero mod p are dropped at construction (src/SpaSM.jl:955-959)
    B = oracle.CSR(np.array([[prime, 1], [2 * prime, 0]]))
    assert B.nnz() == 1


@pytest.mark.parametrize("prime", [7, 42013, 4294967291])
def test_cross_path_same_rref(oracle, prime):
    """sparse-only (GPLU), forced-dense and default paths: equal rank and equal canonical RREF"""
    n, m, k = 90, 100, 4
    A = _mk(oracle, n, m, k, prime, 99)
    ref = None
    for kw in (
        dict(),
        dict(sparsity_threshold=0.0, max_round=1),
        dict(sparsity_threshold=0.0, max_round=1, dense_block_size=7),
        dict(enable_dense=False),
        dict(enable_dense=False, enable_greedy_pivot_search=False),
        dict(max_round=0),
        dict(max_round=0, sparsity_threshold=0.0, dense_block_size=13),
    ):
        fact = oracle.echelonize(A, **kw)
        checks.check_U_structure(oracle, fact)
        checks.check_rank_and_rowspace(oracle, A, fact)
        R = checks.canonical_rref(oracle, fact)
        if ref is None:
            ref = R
        assert np.array_equal(ref, R), kw


@pytest.mark.parametrize("prime", [251, 42013, 2147483647])
def test_solve_and_gesv(oracle, prime):
    n, m, k = 50, 80, 4
    A = _mk(oracle, n, m, k, prime, 5)
    fact = oracle.echelonize(A, L=True)
    checks.check_U_structure(oracle, fact)
    checks.check_rank_and_rowspace(oracle, A, fact)
    Ad, Ld, Ud = checks.dense_of(oracle, A), checks.dense_of(oracle, fact.L), checks.dense_of(oracle, fact.U)
    assert fact.L.shape == (n, fact.r)
    assert np.array_equal(checks.mm(Ld, Ud, prime), Ad), "A != L.U"
    rng = np.random.default_rng(1)
    for t in range(6):
        x0 = rng.integers(0, prime, size=n)
        b = checks.mm(x0, Ad, prime)
        if t >= 4:  # almost surely outside the row space when r < m
            b = rng.integers(0, prime, size=m)
        bb = oracle.CSR(np.zeros((1, 1))).field  # noqa: F841  (field helper not needed)
        xb = oracle.solve(fact, pkg_bal(b, prime))
        inside = synth.dense_rank_mod_p(np.vstack([Ad, b]), prime) == fact.r
        assert (xb is not None) == inside
        if xb is not None:
            assert np.array_equal(checks.mm(xb.astype(np.int64) % prime, Ad, prime), b % prime), "x.A != b"
    # gesv: rows of B = some rows of A and a random row
    import scipy.sparse as sp

    Bd = np.vstack([Ad[3], (Ad[1] + 2 * Ad[2]) % prime, rng.integers(0, prime, size=m)])
    B = oracle.CSR(sp.csc_matrix(Bd.T), prime)
    X, ok = oracle.gesv(fact, B)
    assert ok.tolist()[:2] == [True, True]
    Xd = checks.dense_of(oracle, X)
    for i in range(3):
        if ok[i]:
            assert np.array_equal(checks.mm(Xd[i], Ad, prime), Bd[i] % prime)


def pkg_bal(v, prime):
    return synth.balanced(np.asarray(v), prime)


@pytest.mark.parametrize("prime", [3, 42013])
def test_triangular_solve_contract(oracle, prime):
    """src/SpaSM.jl:694-713: x_b.U + x_a == B[k]; xj stays zero; pattern is xj[top:m]"""
    n, m, k = 40, 60, 4
    A = _mk(oracle, n, m, k, prime, 3)
    fact = oracle.echelonize(A)
    U, qinv = fact.U, fact.qinv.copy()
    Ud, Ad = checks.dense_of(oracle, U), checks.dense_of(oracle, A)
    xj = np.zeros(3 * m, dtype=np.int32)
    x = np.full(m, 12345, dtype=np.int32)  # "does not need to be initialized"
    for row in range(0, n, 7):
        top = oracle.sparse_triangular_solve_row(U, A, row, xj, x, qinv)
        pat = xj[top:m].copy()
        xx = np.zeros(m, dtype=np.int64)
        xx[pat] = x[pat]
        xb = np.zeros(U.n, dtype=np.int64)
        piv = qinv[pat] >= 0
        xb[qinv[pat[piv]]] = xx[pat[piv]]
        xa = xx.copy()
        xa[pat[piv]] = 0
        assert np.array_equal((checks.mm(xb.astype(np.int64) % prime, Ud, prime) + xa) % prime, Ad[row])
        assert not xa.any(), "rows of A are in the row space of U"
        assert not xj[2 * m :].any(), "marks must be left zero (\"it remains OK\", src/SpaSM.jl:700)"
        xj[: 2 * m] = 0  # pattern + DFS stack scratch; the Julia driver refills with zeros too (:740)
    X = oracle.sparse_triangular_solve(fact, A)
    assert X is not None and X.shape == (n, U.n)
    assert np.array_equal(checks.mm(checks.dense_of(oracle, X), Ud, prime), Ad)


def test_thread_count_invariance(oracle, pkg):
    """results do not depend on OMP_NUM_THREADS (normalisation N3)"""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path[:0]=['.','tests']; import __graft_entry__ as e, synth, numpy as np, hashlib;"
        "pk=e.load_package(); o=pk.SpaSM(e.ORACLE_LIB);"
        "p,j,x=synth.random_rows(400,450,4,42013,11); A=o.from_arrays(400,450,p,j,x);"
        "f=o.echelonize(A); K=o.kernel(f); h=hashlib.sha256();"
        "[h.update(a.tobytes()) for a in f.U.arrays()+K.arrays()+(f.qinv,)]; print(f.r,h.hexdigest())"
    )
    outs = set()
    for nt in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        outs.add(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(pkg.PKG_DIR.parent), check=True).stdout.strip())
    assert len(outs) == 1, outs


@pytest.mark.parametrize("kw", [dict(), dict(dense_block_size=16, max_round=0), dict(dense_block_size=16, max_round=0, low_rank_start_weight=1)])
@pytest.mark.parametrize("prime", [7, 42013, 4294967291])
def test_low_rank_mode_oracle(oracle, prime, kw):
    """tall-and-skinny finish: rank and row space are right whatever the weights / block size"""
    n, m, k = 900, 70, 2
    A = _mk(oracle, n, m, k, prime, 77)
    lines = []
    oracle.log(lambda s: lines.append(s) or 0)
    try:
        fact = oracle.echelonize(A, verbose=True, **kw)
    finally:
        oracle.log(None)
    assert any("low-rank" in l for l in lines)
    checks.check_U_structure(oracle, fact)
    checks.check_rank_and_rowspace(oracle, A, fact)
    K = oracle.kernel(fact)
    checks.check_kernel(oracle, A, fact, K)


@pytest.mark.parametrize("case", [(1500, 1500, 5, 42013, 3, {}), (900, 1300, 4, 65521, 12, dict(dense_block_size=64)),
                                  (700, 650, 4, 4294967291, 13, dict(dense_block_size=100)), (1600, 1500, 10, 42013, 6, dict(dense_block_size=300)),
                                  (800, 800, 5, 3, 2, dict(dense_block_size=50))])
def test_blocked_dense_equals_rowwise(oracle, case, monkeypatch):
    """the blocked OpenMP dense tail of the oracle (delayed-reduction products, sub-blocked RREF) against its own
    row-by-row form (one sparse triangular solve per row, textbook Gauss-Jordan): every array of the factor, bit for bit"""
    n, m, k, prime, seed, kw = case
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p, j, x, prime)
    blocked = checks.lu_arrays(oracle.echelonize(A, **kw))
    monkeypatch.setenv("SPASM_ORACLE_ROWWISE", "1")
    rowwise = checks.lu_arrays(oracle.echelonize(A, **kw))
    checks.assert_same(rowwise, blocked, "blocked vs row-wise: ")


MIDTAIL = [
    # n, m, k, prime, seed, planted rank (None: random rows), options — inputs whose dense loop meets a block far below full rank
    (2000, 2000, 3, 42013, 4, 300, dict(dense_block_size=100, sparsity_threshold=0.0, max_round=0)),
    (1200, 1000, 4, 65521, 5, None, dict(dense_block_size=64)),
    (1500, 1400, 3, 4294967291, 8, 200, dict(dense_block_size=50, sparsity_threshold=0.0, max_round=0, low_rank_start_weight=2)),
]


def midtail_input(api, case):
    n, m, k, prime, seed, planted, kw = case
    p, j, x = synth.planted_rank(n, m, planted, 0.5, prime, seed) if planted else synth.random_rows(n, m, k, prime, seed)
    return api.from_arrays(n, m, p, j, x, prime), kw


@pytest.mark.parametrize("case", MIDTAIL)
def test_mid_tail_low_rank_switch_oracle(oracle, case, monkeypatch):
    """SURVEY.md A.7 (spasm_schur_dense_randomized, src/SpaSM.jl:767-769): a dense block with fewer than low_rank_ratio * Sn
    pivots hands the remaining rows to the low-rank mode.  The switch fires, the factor is valid, the blocked and the
    row-by-row dense loops agree bit for bit, and without tall-and-skinny the plain dense loop gives the same canonical RREF."""
    A, kw = midtail_input(oracle, case)
    lines = []
    oracle.log(lambda s: lines.append(s) or 0)
    try:
        fact = oracle.echelonize(A, verbose=True, **kw)
    finally:
        oracle.log(None)
    assert any("switching to low-rank" in l for l in lines)
    checks.check_U_structure(oracle, fact)
    checks.check_rank_and_rowspace(oracle, A, fact)
    blocked = checks.lu_arrays(fact)
    plain = oracle.echelonize(A, enable_tall_and_skinny=False, **kw)
    assert plain.r == fact.r
    if case[3] < (1 << 31):  # (object arithmetic beyond: too slow for what rank + row space already say)
        assert np.array_equal(checks.canonical_rref(oracle, fact), checks.canonical_rref(oracle, plain))
    monkeypatch.setenv("SPASM_ORACLE_ROWWISE", "1")
    checks.assert_same(checks.lu_arrays(oracle.echelonize(A, **kw)), blocked, "blocked vs row-wise: ")


DENSE_L = [
    # n, m, k, prime, seed, options: L = True on inputs whose Schur complement is dense
    (600, 600, 5, 42013, 3, {}),
    (900, 700, 6, 65521, 4, dict(dense_block_size=100)),     # full column rank reached before the rows run out
    (500, 800, 5, 4294967291, 5, dict(dense_block_size=64)),
    (400, 400, 4, 7, 6, dict(dense_block_size=50)),          # tiny field: many cancellations, rank-deficient blocks
]


def check_A_equals_LU(api, A_arrays, fact):
    n, m, p, j, x, prime = A_arrays
    Ld = checks.dense_of(api, fact.L).astype(object)
    Ud = checks.dense_of(api, fact.U).astype(object)
    Ad = synth.csr_to_dense(n, m, p, j, x, prime).astype(object)
    assert np.array_equal(Ld.dot(Ud) % prime, Ad % prime), "A != L.U"
    # L is lower-trapezoidal in the order of Lp: the row that holds the diagonal of column c has nothing in later columns
    Lp = fact.p
    for c in range(fact.r):
        row = Ld[Lp[c]]
        assert row[c] % prime != 0 and not any(v % prime for v in row[c + 1:]), f"L is not triangular at column {c}"


@pytest.mark.parametrize("case", DENSE_L)
def test_dense_tail_with_L_oracle(oracle, case):
    """dense tail with L (prototype spasm_ffpack_LU, src/SpaSM.jl:806-812): A = L.U, L triangular under Lp, same rank and row
    space as without L, and solve() works on it"""
    n, m, k, prime, seed, kw = case
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    A = oracle.from_arrays(n, m, p, j, x, prime)
    lines = []
    oracle.log(lambda s: lines.append(s) or 0)
    try:
        fact = oracle.echelonize(A, verbose=True, L=True, **kw)
    finally:
        oracle.log(None)
    assert any("with L" in l for l in lines), "the dense tail with L did not run"
    assert fact.r == oracle.echelonize(A, **kw).r
    checks.check_U_structure(oracle, fact)
    checks.check_rank_and_rowspace(oracle, A, fact)
    check_A_equals_LU(oracle, (n, m, p, j, x, prime), fact)
    rng = np.random.default_rng(seed)
    Ad = synth.csr_to_dense(n, m, p, j, x, prime).astype(object)
    xv = rng.integers(0, prime, size=n)
    b = synth.balanced(np.array(xv.astype(object).dot(Ad) % prime, dtype=np.int64), prime)
    sol = oracle.solve(fact, b)
    assert sol is not None
    got = np.array(np.asarray(sol).astype(object).dot(Ad) % prime, dtype=np.int64)
    assert np.array_equal(got, np.array(b.astype(object) % prime, dtype=np.int64))
