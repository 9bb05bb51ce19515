"""SMS text format (src/SpaSM.jl:498-529, :1029-1086): "n m M" header, 1-based "i j v" lines, "0 0 0" terminator.
Round trips through the host code of the product library and of the oracle, a fixture written the way the reference's
own writer does it (src/SpaSM.jl:531-545), and — on the GPU — an echelonization of a matrix that came from a file."""
from pathlib import Path

import numpy as np
import pytest

import checks
import synth

ROOT = Path(__file__).resolve().parent.parent
FIXTURE = ROOT / "tests" / "golden" / "banded_60x70.sms"


def _same(A, B):
    assert A.shape == B.shape and A.prime == B.prime
    for a, b in zip(A.arrays(), B.arrays()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_sms_round_trip(pkg, oracle, product_lib, which, tmp_path):
    api = oracle if which == "oracle" else pkg.SpaSM(product_lib)  # SMS I/O is host code: it works without a GPU
    n, m, prime = 37, 53, 42013
    p, j, x = synth.ragged_rows(n, m, 4, prime, 11)
    A = api.from_arrays(n, m, p, j, x, prime)
    f = tmp_path / "a.sms"
    api.save(f, A)
    text = f.read_text().splitlines()
    assert text[0] == f"{n} {m} M" and text[-1] == "0 0 0" and len(text) == A.nnz() + 2
    i0, j0, v0 = map(int, text[1].split())
    assert i0 >= 1 and j0 >= 1 and -prime // 2 <= v0 <= prime // 2  # 1-based, balanced representatives
    B = api.load(f, prime)
    # spasm_compress keeps the entries of a row in file order: the round trip is the identity on (p, j, x)
    _same(A, B)


def test_sms_cross_library(pkg, oracle, product_lib, tmp_path):
    prod = pkg.SpaSM(product_lib)
    A = oracle.CSR(np.array([[1, 2, 0, 0], [0, 0, 0, 0], [0, 0, 3, -4]]))
    f = tmp_path / "x.sms"
    oracle.save(f, A)
    _same(A, prod.load(f))
    prod.save(f, prod.load(f))
    _same(A, oracle.load(f))


def test_sms_fixture_as_the_reference_writes_it(oracle):
    """tests/golden/banded_60x70.sms was written by tests/golden/make_sms_fixture.py following the reference's pure-Julia
    writer (src/SpaSM.jl:531-545): entries in COLUMN-major order of the SparseMatrixCSC, values as Int(v)"""
    A = oracle.load(FIXTURE)
    assert A.shape == (60, 70) and A.nnz() == 289
    fact = oracle.echelonize(A)
    checks.check_U_structure(oracle, fact)
    checks.check_rank_and_rowspace(oracle, A, fact)


@pytest.mark.gpu
def test_sms_file_echelonize_on_gpu(gpu, oracle):
    A = gpu.load(FIXTURE)
    checks.assert_same(checks.lu_arrays(oracle.echelonize(oracle.load(FIXTURE))), checks.lu_arrays(gpu.echelonize(A)), "sms: ")
