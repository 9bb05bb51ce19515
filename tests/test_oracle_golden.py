"""Pin the oracle against every golden vector the reference holds for this path:
test/runtests.jl:7-24 (construction, transpose round trip, two kernels of a 3x4 matrix mod 42013)
and the README transcript (README.md:9-48).  Committed as tests/golden/reference_goldens.json
(made by tests/golden/make_goldens.py from the literal values in those files)."""
import json
from pathlib import Path

import numpy as np
import scipy.sparse as sp

G = json.loads((Path(__file__).parent / "golden" / "reference_goldens.json").read_text())


def _julia_sparse(I, J, V, shape=None):
    I, J = np.array(I) - 1, np.array(J) - 1
    if shape is None:
        shape = (I.max() + 1, J.max() + 1)
    return sp.csc_matrix((np.array(V, dtype=np.int64), (I, J)), shape=shape)


def _eq_mod(a, b, p):
    a, b = sp.csc_matrix(a), sp.csc_matrix(b)
    if a.shape != b.shape:
        return False
    d = (a - b).toarray() % p
    return not d.any()


def test_construction(oracle):
    g = G["runtests"]
    m = _julia_sparse(g["I"], g["J"], g["V"])
    sm = oracle.CSR(m)
    assert sm.shape == (4, 3)  # a column of m is a SpaSM row (src/SpaSM.jl:941-968, README.md:7)
    assert _eq_mod(oracle.sparse(sm), m, g["prime"])
    assert (oracle.sparse(sm) != sp.csc_matrix(m)).nnz == 0  # small values: representation is identical


def test_transpose_roundtrip(oracle):
    g = G["runtests"]
    sm = oracle.CSR(_julia_sparse(g["I"], g["J"], g["V"]))
    tt = oracle.transpose(oracle.transpose(sm))
    assert (oracle.sparse(tt) != oracle.sparse(sm)).nnz == 0
    t = oracle.transpose(sm)
    assert t.shape == (3, 4)
    assert (oracle.sparse(t) != sp.csc_matrix(oracle.sparse(sm).T)).nnz == 0


def test_kernel_goldens(oracle):
    g = G["runtests"]
    p = g["prime"]
    sm = oracle.CSR(_julia_sparse(g["I"], g["J"], g["V"]))
    k = oracle.kernel(sm)
    k1 = g["kernel"]
    assert _eq_mod(oracle.sparse(k), _julia_sparse(k1["I"], k1["J"], k1["V"], tuple(k1["shape"])), p)
    k2 = g["kernel_transpose"]
    kt = oracle.kernel(oracle.transpose(sm))
    assert _eq_mod(oracle.sparse(kt), _julia_sparse(k2["I"], k2["J"], k2["V"], tuple(k2["shape"])), p)
    # exact balanced representation: 42012 == -1, 28010 stays
    F = oracle.sparse(kt).toarray()
    assert sorted(F[F != 0].tolist()) == sorted([2, -1, -14003 if 28010 > p // 2 else 28010, -1])


def test_readme_transcript(oracle):
    g = G["readme"]
    p = g["prime"]
    lines = []
    oracle.log(lambda s: lines.append(s) or 0)
    try:
        sm = oracle.CSR(_julia_sparse(g["I"], g["J"], g["V"]))
        assert repr(sm) == "2x2 CSR matrix % 42013 with 4 (maximum 4) non-zeros"
        fact = oracle.echelonize(sm, verbose=True)
        k = oracle.kernel(fact, verbose=True)
    finally:
        oracle.log(None)
    assert fact.r == g["rank"]
    assert fact.U.nnz() == g["nz_in_basis"]
    assert k.nnz() == g["nnz_K"]
    kk = g["kernel"]
    assert _eq_mod(oracle.sparse(k), _julia_sparse(kk["I"], kk["J"], kk["V"], tuple(kk["shape"])), p)
    text = "".join(lines)
    for must in g["log_must_contain"]:
        assert must in text, must
