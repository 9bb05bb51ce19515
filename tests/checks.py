"""Canonical invariants (SURVEY.md §8c level 3) — independent of any restatement of libspasm.
Used for the oracle on CPU and for the CUDA library on the GPU box."""
from __future__ import annotations

import numpy as np
import synth


def mm(A, B, prime):
    """exact (A @ B) mod prime for residues in [0, prime)"""
    A, B = np.asarray(A), np.asarray(B)
    if prime * prime * max(A.shape[-1], 1) < (1 << 62):
        return (A.astype(np.int64) @ B.astype(np.int64)) % prime
    return (A.astype(object) @ B.astype(object) % prime).astype(np.int64)


def dense_of(api, M):
    p, j, x = M.arrays()
    return synth.csr_to_dense(M.n, M.m, p, j, x, M.prime)


def check_U_structure(api, fact):
    """unit pivot stored first, qinv consistent, each U row only references pivots of later rows"""
    U, qinv = fact.U, fact.qinv
    p, j, x = U.arrays()
    r = fact.r
    assert U.n == r and len(p) == r + 1
    assert (np.diff(p) >= 1).all()
    heads = j[p[:-1]]
    assert (x[p[:-1]] == 1).all(), "pivot entries must be 1"
    assert (qinv[heads] == np.arange(r)).all(), "qinv[pivot col] != row"
    assert (qinv >= 0).sum() == r
    rows = np.repeat(np.arange(r), np.diff(p))
    ref = qinv[j]
    bad = (ref >= 0) & (ref < rows)
    assert not bad.any(), "U row references an earlier pivot: not upper triangular in row order"
    # distinct columns inside each row
    key = rows.astype(np.int64) * U.m + j
    assert len(np.unique(key)) == len(key)
    assert (x != 0).all(), "explicit zero stored in U"


def check_rank_and_rowspace(api, A, fact):
    prime = A.prime
    Ad = dense_of(api, A)
    Ud = dense_of(api, fact.U)
    r0 = synth.dense_rank_mod_p(Ad, prime)
    assert fact.r == r0, f"rank {fact.r} != dense Gauss rank {r0}"
    assert synth.dense_rank_mod_p(np.vstack([Ad, Ud]), prime) == r0, "rowspace(U) != rowspace(A)"


def check_kernel(api, A, fact, K):
    prime = A.prime
    Ad, Kd = dense_of(api, A), dense_of(api, K)
    assert K.shape == (A.m - fact.r, A.m)
    assert not mm(Ad, Kd.T, prime).any(), "A.K^T != 0"
    assert synth.dense_rank_mod_p(Kd, prime) == A.m - fact.r, "kernel basis is not independent"
    # rows in increasing free column, leading entry (j,-1)
    p, j, x = K.arrays()
    free = np.nonzero(fact.qinv < 0)[0]
    assert (j[p[:-1]] == free).all()
    assert (x[p[:-1]] == -1).all()


def check_rref(api, fact, R, Rqinv):
    prime = fact.U.prime
    Rd = dense_of(api, R)
    Ud = dense_of(api, fact.U)
    r = fact.r
    assert R.shape == (r, fact.U.m)
    pivcols = np.nonzero(fact.qinv >= 0)[0]
    sub = Rd[:, pivcols]
    assert ((sub != 0).sum(axis=0) == 1).all() and ((sub != 0).sum(axis=1) == 1).all(), "R is not reduced"
    assert (Rqinv[pivcols] == fact.qinv[pivcols]).all()
    assert synth.dense_rank_mod_p(np.vstack([Rd, Ud]), prime) == r


def canonical_rref(api, fact):
    """reduced row echelon form of rowspace(U) as a dense matrix — canonical for the row space"""
    prime = fact.U.prime
    Ud = dense_of(api, fact.U).astype(np.int64)
    D = Ud.copy() if prime <= (1 << 31) else Ud.astype(object)
    n, m = D.shape
    r = 0
    for c in range(m):
        if r == n:
            break
        nz = np.nonzero(D[r:, c])[0]
        if len(nz) == 0:
            continue
        piv = r + int(nz[0])
        D[[r, piv]] = D[[piv, r]]
        D[r] = D[r] * pow(int(D[r, c]), -1, prime) % prime
        rows = np.nonzero(D[:, c])[0]
        rows = rows[rows != r]
        D[rows] = (D[rows] - np.outer(D[rows, c], D[r])) % prime
        r += 1
    return D[:r].astype(np.int64)


def lu_arrays(fact):
    """everything that must be bit-exact between two libraries"""
    out = {"r": fact.r, "qinv": fact.qinv.copy()}
    out["Up"], out["Uj"], out["Ux"] = fact.U.arrays()
    if fact.data.contents.L:
        out["Lp"], out["Lj"], out["Lx"] = fact.L.arrays()
        out["p"] = fact.p.copy()
    return out


def assert_same(a: dict, b: dict, what=""):
    assert a.keys() == b.keys(), (a.keys(), b.keys())
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert a[k].shape == b[k].shape, f"{what}{k}: shape {a[k].shape} vs {b[k].shape}"
            if not np.array_equal(a[k], b[k]):
                bad = np.nonzero(a[k] != b[k])[0]
                raise AssertionError(f"{what}{k}: {len(bad)} mismatches, first at {bad[:5]}: {a[k][bad[:5]]} vs {b[k][bad[:5]]}")
        else:
            assert a[k] == b[k], f"{what}{k}: {a[k]} vs {b[k]}"
