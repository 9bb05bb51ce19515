"""CPU-side checks of the drop-in boundary: the CUDA library loads without a GPU, exports every
symbol include/spasm_b200.h declares, mirrors the reference's struct layouts, and
refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "spasm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spasm_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(pkg, product_lib):
    lib = C.CDLL(str(product_lib), mode=C.RTLD_LOCAL)
    syms = _declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/spasm_b200.h but not exported"
    C.c_void_p.in_dll(lib, "logcallback")  # data symbol (src/SpaSM.jl:35)
    assert set(pkg.ABI) == set(syms), set(pkg.ABI) ^ set(syms)


def test_extension_header_symbols_exported(product_lib):
    """include/spasm_b200_ext.h: the entry points libspasm does not have (memory policy, multi-GPU, instrumentation)"""
    text = (ROOT / "include" / "spasm_b200_ext.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    syms = sorted(set(re.findall(r"\b(spasm_b200_[A-Za-z0-9_]+)\s*\(", text)))
    assert len(syms) >= 20
    lib = C.CDLL(str(product_lib), mode=C.RTLD_LOCAL)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/spasm_b200_ext.h but not exported"
    # and nothing is exported that no header declares
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", str(product_lib)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T spasm_" in l}
    declared = set(syms) | set(_declared_symbols())
    assert exported <= declared, exported - declared


def test_oracle_exports_same_abi(pkg, oracle):
    for s in _declared_symbols():
        assert hasattr(oracle.lib, s)


def test_struct_layouts(pkg):
    # numbers from SURVEY.md §8b (gcc sizeof/offsetof on the C mirrors of the Julia structs)
    assert C.sizeof(pkg._Field) == 32
    assert C.sizeof(pkg._CSR) == 72 and pkg._CSR.p.offset == 16 and pkg._CSR.j.offset == 24 and pkg._CSR.x.offset == 32 and pkg._CSR.field.offset == 40
    assert C.sizeof(pkg._Triplet) == 80
    assert C.sizeof(pkg._LU) == 48 and pkg._LU.L.offset == 8 and pkg._LU.U.offset == 16 and pkg._LU.qinv.offset == 24 and pkg._LU.p.offset == 32
    O = pkg.EchelonizeOpts
    assert C.sizeof(O) == 64
    assert (O.min_pivot_proportion.offset, O.max_round.offset, O.sparsity_threshold.offset, O.dense_block_size.offset) == (8, 16, 24, 32)
    assert (O.low_rank_ratio.offset, O.tall_and_skinny_ratio.offset, O.low_rank_start_weight.offset) == (40, 48, 56)


def test_host_containers_without_gpu(pkg, product_lib):
    """allocation / triplet / opts / scalar field ops are host code and work without a device"""
    api = pkg.SpaSM(product_lib)
    assert api.backend == "cuda-sm_100a"
    o = api.EchelonizeOpts()
    assert (o.enable_greedy_pivot_search, o.enable_tall_and_skinny, o.enable_dense, o.enable_GPLU, o.L, o.complete) == (True, True, True, True, False, False)
    assert (o.min_pivot_proportion, o.max_round, o.sparsity_threshold, o.dense_block_size) == (0.1, 3, 0.05, 1000)
    A = api.CSR(np.array([[1, 2, 0, 0], [0, 0, 0, 0], [0, 0, 3, 4]]))
    assert A.shape == (4, 3) and A.nnz() == 4
    assert repr(A) == "4x3 CSR matrix % 42013 with 4 (maximum 4) non-zeros"
    T = api.lib.spasm_triplet_alloc(0, 0, 1, 42013, True)
    for (i, j, v) in [(0, 1, 5), (2, 0, -1), (0, 1, 42013 - 5), (1, 1, 7), (1, 1, 1)]:
        api.lib.spasm_add_entry(T, i, j, v)
    M = pkg.CSR(api, api.lib.spasm_compress(T))
    api.lib.spasm_triplet_free(T)
    assert M.shape == (3, 2)
    assert api.sparse(M).T.toarray().tolist() == [[0, 0], [0, 8], [-1, 0]]  # 5 + (-5) cancels
    F = pkg._Field()
    api.lib.spasm_field_init(42013, C.byref(F))
    assert api.lib.spasm_ZZp_mul(C.byref(F), 21006, 21006) == pkg.ZZp(42013, 21006 * 21006).v
    assert api.lib.spasm_ZZp_inverse(C.byref(F), 3) == 28009 - 42013  # README / SURVEY §4: inv(3) = 28009


def test_no_cpu_fallback(pkg, product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    api = pkg.SpaSM(product_lib)
    msgs = []
    api.log(lambda s: msgs.append(s) or 0)
    A = api.CSR(np.array([[1, 2], [3, 6]]))
    with pytest.raises(RuntimeError):
        api.echelonize(A)
    with pytest.raises(RuntimeError):
        api.transpose(A)
    api.log(None)
    assert any("no CUDA device" in m for m in msgs)


def test_missing_library_is_loud(pkg, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.SpaSM(tmp_path / "libspasm_b200.so")
