"""Seeded synthetic inputs shared by tests/, bench.py and __graft_entry__.smoke().

Language-neutral counter-based PRNG (splitmix64 of seed + counter) so that C, Python and Julia
hosts can regenerate identical matrices.  All generators return raw CSR arrays
(p int64[n+1], j int32[nnz], x int32[nnz] balanced) of the *SpaSM* matrix (rows = SpaSM rows).
"""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z: np.ndarray) -> np.ndarray:
    """vectorised finaliser of splitmix64 applied to counters z (uint64)"""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
        return z ^ (z >> np.uint64(31))


def stream(seed: int, lo: int, hi: int) -> np.ndarray:
    """draws number lo..hi-1 of the stream `seed`"""
    with np.errstate(over="ignore"):
        ctr = np.arange(lo, hi, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    return splitmix64(ctr)


def balanced(v: np.ndarray, prime: int) -> np.ndarray:
    v = np.mod(v.astype(np.int64), prime)
    return np.where(v > prime // 2, v - prime, v).astype(np.int32)


def random_rows(n: int, m: int, k: int, prime: int, seed: int, sort_cols: bool = False):
    """n x m, exactly k distinct uniformly random columns per row (k <= m), values uniform in
    [1, p-1].  Columns are kept in draw order (storage order matters to the pivot search)."""
    assert k <= m
    extra = 4
    draws = (stream(seed, 0, n * (k + extra)) % np.uint64(m)).astype(np.int64).reshape(n, k + extra)
    cols = np.empty((n, k), dtype=np.int32)
    # fast path: first k draws distinct
    first = draws[:, :k]
    srt = np.sort(first, axis=1)
    ok = (np.diff(srt, axis=1) != 0).all(axis=1) if k > 1 else np.ones(n, dtype=bool)
    cols[ok] = first[ok]
    for i in np.nonzero(~ok)[0]:
        seen, out, t = set(), [], 0
        row = draws[i]
        while len(out) < k:
            c = int(row[t]) if t < k + extra else int(stream(seed ^ 0xABCDEF, i * 1000 + t, i * 1000 + t + 1)[0] % np.uint64(m))
            t += 1
            if c not in seen:
                seen.add(c)
                out.append(c)
        cols[i] = out
    if sort_cols:
        cols.sort(axis=1)
    vals = (stream(seed + 1, 0, n * k) % np.uint64(prime - 1)).astype(np.int64) + 1
    p = np.arange(0, (n + 1) * k, k, dtype=np.int64)
    return p, cols.reshape(-1), balanced(vals, prime)


def ragged_rows(n: int, m: int, kmax: int, prime: int, seed: int):
    """rows of random length 0..kmax (empty rows included), distinct columns, draw order"""
    lens = (stream(seed + 2, 0, n) % np.uint64(kmax + 1)).astype(np.int64)
    lens = np.minimum(lens, m)
    p = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    j = np.empty(int(p[-1]), dtype=np.int32)
    rng = np.random.Generator(np.random.PCG64(seed))
    for i in range(n):
        j[p[i] : p[i + 1]] = rng.choice(m, size=int(lens[i]), replace=False)
    vals = (stream(seed + 1, 0, int(p[-1])) % np.uint64(prime - 1)).astype(np.int64) + 1
    return p, j, balanced(vals, prime)


def planted_rank(n: int, m: int, r: int, extra: float, prime: int, seed: int, combo: int = 3):
    """GL7d19-shaped planted-rank matrix.

    r "basis" rows: row b has a leading entry on staircase column s_b (distinct, increasing) and
    ~Poisson(extra) further entries to the right, values in {+-1,+-2,+-3}; the other n-r rows are
    combinations of 2..combo random basis rows with small coefficients.  Rows and columns are then
    permuted by seeded permutations.  rank == r by construction (the basis rows are in echelon
    form)."""
    assert r <= min(n, m)
    rng = np.random.Generator(np.random.PCG64(seed))
    stair = np.sort(rng.choice(m, size=r, replace=False)).astype(np.int64)
    small = np.array([1, -1, 2, -2, 3, -3], dtype=np.int64)
    cnt = rng.poisson(extra, size=r)
    room = (m - 1 - stair).astype(np.int64)
    cnt = np.minimum(cnt, room)
    # basis rows as (row, col, val) triples
    tot = int(cnt.sum())
    rows_b = np.repeat(np.arange(r), cnt)
    offs = rng.random(tot)
    cols_b = stair[rows_b] + 1 + np.floor(offs * room[rows_b]).astype(np.int64)
    vals_b = small[rng.integers(0, 6, size=tot)]
    rows_all = np.concatenate([np.arange(r), rows_b])
    cols_all = np.concatenate([stair, cols_b])
    vals_all = np.concatenate([small[rng.integers(0, 6, size=r)], vals_b])
    import scipy.sparse as sp

    B = sp.csr_matrix((vals_all, (rows_all, cols_all)), shape=(r, m))
    B.sum_duplicates()
    nd = n - r
    if nd > 0:
        k = rng.integers(2, combo + 1, size=nd)
        rr = np.repeat(np.arange(nd), k)
        cc = rng.integers(0, r, size=int(k.sum()))
        vv = small[rng.integers(0, 4, size=int(k.sum()))]
        Cmat = sp.csr_matrix((vv, (rr, cc)), shape=(nd, r))
        Cmat.sum_duplicates()
        D = (Cmat @ B).tocsr()
        A = sp.vstack([B, D]).tocsr()
    else:
        A = B
    A.data = np.mod(A.data, prime)
    A.eliminate_zeros()
    rp = rng.permutation(n)
    cp = rng.permutation(m)
    A = A[rp][:, cp].tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), balanced(A.data, prime)


def dense_random(n: int, m: int, prime: int, seed: int, rank: int | None = None) -> np.ndarray:
    """row-major int32 balanced dense matrix; optional planted rank (product of two factors)"""
    if rank is None:
        v = (stream(seed, 0, n * m) % np.uint64(prime)).astype(np.int64).reshape(n, m)
        return balanced(v, prime).reshape(n, m)
    a = (stream(seed, 0, n * rank) % np.uint64(prime)).astype(np.int64).reshape(n, rank)
    b = (stream(seed + 7, 0, rank * m) % np.uint64(prime)).astype(np.int64).reshape(rank, m)
    if prime < (1 << 28):
        out = np.zeros((n, m), dtype=np.int64)
        for k0 in range(0, rank, 64):  # keep sums of products below 2^63
            out = (out + a[:, k0 : k0 + 64] @ b[k0 : k0 + 64]) % prime
    else:  # exact Python integers
        out = ((a.astype(object) @ b.astype(object)) % prime).astype(np.int64)
    return balanced(out, prime).reshape(n, m)


# ----------------------------------------------------------------------------- checkers (numpy, small sizes)
def dense_rank_mod_p(A: np.ndarray, prime: int) -> int:
    """plain Gauss elimination mod p on an int64 copy — independent of any restatement"""
    A = np.mod(np.array(A, dtype=np.int64), prime)
    if prime > (1 << 31):  # products overflow int64: exact Python integers
        A = A.astype(object)
    n, m = A.shape
    r = 0
    for c in range(m):
        if r == n:
            break
        nzr = np.nonzero(A[r:, c])[0]
        if len(nzr) == 0:
            continue
        piv = r + int(nzr[0])
        if piv != r:
            A[[r, piv]] = A[[piv, r]]
        inv = pow(int(A[r, c]), -1, prime)
        A[r] = (A[r] * inv) % prime
        rows = np.nonzero(A[:, c])[0]
        rows = rows[rows != r]
        if len(rows):
            A[rows] = (A[rows] - np.outer(A[rows, c], A[r])) % prime
        r += 1
    return r


def csr_to_dense(n, m, p, j, x, prime) -> np.ndarray:
    D = np.zeros((n, m), dtype=np.int64)
    rows = np.repeat(np.arange(n), np.diff(p))
    np.add.at(D, (rows, j[: p[-1]]), x[: p[-1]].astype(np.int64))
    return np.mod(D, prime)


def banded_planted(n: int, m: int, r: int, extra: float, window: int, prime: int, seed: int, combo: int = 3, spread: int = 64, colblock: int = 8):
    """GL7d19-shaped planted-rank matrix that STAYS SPARSE under elimination.

    r basis rows in echelon form: row b has a leading entry on staircase column s_b and
    ~Poisson(extra) further entries at LOCAL offsets (exponential, mean `window`) to the right;
    the other n-r rows are combinations of 2..combo basis rows that are close to each other
    (indices within `spread`).  Rows are shuffled globally, columns only inside blocks of
    `colblock` (so leftmost-entry structure is perturbed but locality survives).  rank == r."""
    assert r <= min(n, m)
    import scipy.sparse as sp

    rng = np.random.Generator(np.random.PCG64(seed))
    stair = np.sort(rng.choice(m, size=r, replace=False)).astype(np.int64)
    small = np.array([1, -1, 2, -2, 3, -3], dtype=np.int64)
    cnt = rng.poisson(extra, size=r)
    rows_b = np.repeat(np.arange(r), cnt)
    offs = 1 + np.floor(rng.exponential(window, size=len(rows_b))).astype(np.int64)
    cols_b = stair[rows_b] + offs
    keep = cols_b < m
    rows_b, cols_b = rows_b[keep], cols_b[keep]
    vals_b = small[rng.integers(0, 6, size=len(rows_b))]
    B = sp.csr_matrix((np.concatenate([small[rng.integers(0, 6, size=r)], vals_b]),
                       (np.concatenate([np.arange(r), rows_b]), np.concatenate([stair, cols_b]))), shape=(r, m))
    B.sum_duplicates()
    nd = n - r
    if nd > 0:
        k = rng.integers(2, combo + 1, size=nd)
        rr = np.repeat(np.arange(nd), k)
        centre = np.repeat(rng.integers(0, r, size=nd), k)
        cc = np.clip(centre + rng.integers(-spread, spread + 1, size=len(rr)), 0, r - 1)
        vv = small[rng.integers(0, 4, size=len(rr))]
        Cm = sp.csr_matrix((vv, (rr, cc)), shape=(nd, r))
        Cm.sum_duplicates()
        A = sp.vstack([B, (Cm @ B).tocsr()]).tocsr()
    else:
        A = B
    A.data = np.mod(A.data, prime)
    A.eliminate_zeros()
    rp = rng.permutation(n)
    cp = np.arange(m)
    for c0 in range(0, m, colblock):  # local column shuffle
        pass
    nblk = (m + colblock - 1) // colblock
    key = np.repeat(np.arange(nblk), colblock)[:m].astype(np.float64) + rng.random(m)
    cp = np.argsort(key, kind="stable")
    A = A[rp][:, cp].tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), balanced(A.data, prime)


def sparse_chain_planted(n: int, m: int, r: int, extra: float, link: float, window: int, prime: int, seed: int, combo: int = 3, spread: int = 64):
    """GL7d19-shaped planted-rank matrix with SHORT elimination chains (stays sparse at any scale).

    Columns are split into r staircase columns (one per basis row) and m-r free columns.  Basis row b
    has its leading entry on its staircase column, ~Poisson(extra) entries on FREE columns within a
    local window and ~Poisson(link) < 1 entries on later staircase columns, so the pivot graph is
    sub-critical: every row reaches O(1/(1-link)) pivot rows whatever the size.  The other n-r rows are
    combinations of 2..combo nearby basis rows.  Rows are shuffled globally.  rank == r."""
    assert r <= min(n, m) and link < 1.0
    import scipy.sparse as sp

    rng = np.random.Generator(np.random.PCG64(seed))
    is_stair = np.zeros(m, dtype=bool)
    stair = np.sort(rng.choice(m, size=r, replace=False)).astype(np.int64)
    is_stair[stair] = True
    free = np.nonzero(~is_stair)[0].astype(np.int64)
    nf = len(free)
    # position of each staircase column among the free columns (for local windows)
    fpos = np.searchsorted(free, stair)
    small = np.array([1, -1, 2, -2, 3, -3], dtype=np.int64)
    cnt = rng.poisson(extra, size=r)
    rows_f = np.repeat(np.arange(r), cnt)
    idx = np.clip(fpos[rows_f] + rng.integers(-window, window + 1, size=len(rows_f)), 0, max(nf - 1, 0))
    cols_f = free[idx] if nf else np.zeros(0, dtype=np.int64)
    lk = rng.poisson(link, size=r)
    rows_l = np.repeat(np.arange(r), lk)
    tgt = rows_l + 1 + rng.integers(0, window, size=len(rows_l))
    ok = tgt < r
    rows_l, cols_l = rows_l[ok], stair[tgt[ok]]
    rows_all = np.concatenate([np.arange(r), rows_f, rows_l])
    cols_all = np.concatenate([stair, cols_f, cols_l])
    vals_all = small[rng.integers(0, 6, size=len(rows_all))]
    B = sp.csr_matrix((vals_all, (rows_all, cols_all)), shape=(r, m))
    B.sum_duplicates()
    B.data[B.data == 0] = 1
    nd = n - r
    if nd > 0:
        k = rng.integers(2, combo + 1, size=nd)
        rr = np.repeat(np.arange(nd), k)
        centre = np.repeat(rng.integers(0, r, size=nd), k)
        cc = np.clip(centre + rng.integers(-spread, spread + 1, size=len(rr)), 0, r - 1)
        vv = small[rng.integers(0, 4, size=len(rr))]
        Cm = sp.csr_matrix((vv, (rr, cc)), shape=(nd, r))
        Cm.sum_duplicates()
        A = sp.vstack([B, (Cm @ B).tocsr()]).tocsr()
    else:
        A = B
    A.data = np.mod(A.data, prime)
    A.eliminate_zeros()
    A = A[rng.permutation(n)].tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), balanced(A.data, prime)
