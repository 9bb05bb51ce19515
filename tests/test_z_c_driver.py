"""The C driver (tests/c_driver/driver.c) talks to the library through dlopen/dlsym only — the way a C or Julia
host would — and replays the reference's own test vectors (test/runtests.jl:7-24, README.md:9-48).
Kernel bases are compared as the reference does: sparse(k) is the TRANSPOSE of the SpaSM matrix."""
import json
import subprocess
from pathlib import Path

import pytest

import __graft_entry__ as entry

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_goldens.json").read_text())


def build_driver(tmp_path):
    exe = tmp_path / "c_driver"
    subprocess.run(["/usr/bin/gcc", "-O1", "-std=gnu11", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c_driver" / "driver.c"), "-ldl", "-o", str(exe)],
                   check=True)
    return exe


def run_case(exe, lib, case):
    r = subprocess.run([str(exe), str(lib), case], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    head = {l.split()[0]: [int(t) for t in l.split()[1:]] for l in lines[:3]}
    entries = sorted(tuple(int(t) for t in l.split()) for l in lines[3:])
    return head, entries


def golden_entries(k):
    # Julia (I, J, V), 1-based, of sparse(kernel) = transpose of the SpaSM matrix -> SpaSM (row, col, v) 0-based
    return sorted((j - 1, i - 1, v) for i, j, v in zip(k["I"], k["J"], k["V"]))


def check(exe, lib):
    head, ent = run_case(exe, lib, "runtests")
    g = GOLD["runtests"]["kernel"]
    assert head["kernel"][:2] == [g["shape"][1], g["shape"][0]] and ent == golden_entries(g)
    head, ent = run_case(exe, lib, "runtests_t")
    g = GOLD["runtests"]["kernel_transpose"]
    assert head["kernel"][:2] == [g["shape"][1], g["shape"][0]] and ent == golden_entries(g)
    head, ent = run_case(exe, lib, "readme")
    g = GOLD["readme"]
    assert head["rank"] == [g["rank"]] and head["nnzU"] == [g["nz_in_basis"]] and head["kernel"][2] == g["nnz_K"]
    assert ent == golden_entries(g["kernel"])


def test_c_driver_on_oracle(tmp_path):
    check(build_driver(tmp_path), entry.build_oracle())


@pytest.mark.gpu
def test_c_driver_on_cuda_library(tmp_path):
    check(build_driver(tmp_path), entry.build_product())


def test_c_driver_product_refuses_without_gpu(tmp_path):
    """on a box without a GPU the CUDA library must load, export the ABI and fail loudly — never compute on the CPU"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(build_driver(tmp_path)), str(entry.build_product()), "readme"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 4 and "no CPU fallback" in r.stderr and "rank" not in r.stdout


@pytest.mark.gpu
def test_degenerate_inputs_on_cuda_library(gpu, oracle):
    """all-zero matrices, a single entry, the identity (tests/test_oracle_invariants.py::test_edge_cases on the GPU)"""
    import numpy as np

    import checks

    for n, m in [(5, 7), (1, 1), (7, 5)]:
        Z = gpu.spzeros(gpu.CSR(np.zeros((1, 1))).field, n, m)
        fact = gpu.echelonize(Z)
        assert fact.r == 0
        K = gpu.kernel(fact)
        assert K.shape == (m, m) and K.nnz() == m
    for M in (np.array([[0, 0, 5], [0, 0, 0]]), np.eye(6, dtype=np.int64), np.ones((4, 9), dtype=np.int64)):
        A = gpu.CSR(M)
        fo, fg = oracle.echelonize(A), gpu.echelonize(A)
        checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(fg))
        for a, b in zip(oracle.kernel(fo).arrays(), gpu.kernel(fg).arrays()):
            assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("lazy_k", ["64", "96", "0"])
def test_deferred_updates_small_depth_bit_exact(gpu, oracle, lazy_k, monkeypatch):
    """the deferred trailing updates with a tiny flush depth (near rows = a few 16-row panels, many corrections and
    flushes) against the oracle, bit for bit.  At the default depth the far path only runs on matrices too large
    for the oracle; there it is covered by the planted-rank and rank(A) = rank(A^T) tests."""
    import checks
    import synth

    import ctypes as C

    monkeypatch.setenv("SPASM_B200_LAZY_K", lazy_k)
    st = (C.c_longlong * 4)()
    gpu.lib.spasm_b200_tail_stats.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    for (n, m, k, prime, seed) in [(1500, 1500, 5, 42013, 3), (900, 1300, 4, 65521, 12), (700, 650, 4, 4294967291, 13)]:
        p, j, x = synth.random_rows(n, m, k, prime, seed)
        A = gpu.from_arrays(n, m, p, j, x, prime)
        kw = dict(dense_block_size=16, sparsity_threshold=0.0, max_round=1)
        gpu.lib.spasm_b200_tail_stats(st, 1)
        got = checks.lu_arrays(gpu.echelonize(A, **kw))
        gpu.lib.spasm_b200_tail_stats(st, 0)
        if lazy_k != "0":  # the far-row path really ran: flushes with far rows and multiplier corrections
            assert st[0] > 0 and st[1] > 0, list(st)
        else:
            assert st[0] == 0 and st[1] == 0, list(st)
        checks.assert_same(checks.lu_arrays(oracle.echelonize(A, **kw)), got, f"lazy_k={lazy_k}: ")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(9000, 9000, 10, 42013, 6, {}, None), (9000, 8000, 12, 65521, 21, dict(dense_block_size=500), None),
                                  (5200, 5000, 9, 2147483647, 22, dict(dense_block_size=300), "1200")])
def test_deferred_updates_default_depth_bit_exact(gpu, oracle, case, monkeypatch):
    """the bench's own code path: DEFAULT flush depth (4096), more than 4096 rows in the dense tail, so the far rows
    receive deferred products and corrected multipliers — every entry of U against the oracle, bit for bit"""
    import ctypes as C

    import checks
    import synth

    n, m, k, prime, seed, kw, lazy_k = case
    if lazy_k is not None:  # 31-bit prime (CUDA-core / multi-limb GEMM): a tail the oracle finishes quickly, depth 1200
        monkeypatch.setenv("SPASM_B200_LAZY_K", lazy_k)
    p, j, x = synth.random_rows(n, m, k, prime, seed)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    st = (C.c_longlong * 4)()
    gpu.lib.spasm_b200_tail_stats.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    gpu.lib.spasm_b200_tail_stats(st, 1)
    got = checks.lu_arrays(gpu.echelonize(A, **kw))
    gpu.lib.spasm_b200_tail_stats(st, 0)
    assert st[0] > 0 and st[1] > 0 and st[2] > 0, f"the far-row path did not run: {list(st)}"
    checks.assert_same(checks.lu_arrays(oracle.echelonize(A, **kw)), got, "default depth: ")


@pytest.mark.gpu
def test_library_returns_device_memory(gpu):
    """after an entry point returns the library holds (almost) no device memory: a second process or CUDA.jl in the
    same session must find the HBM free (runtime.cu memory policy); spasm_b200_set_cache(1) opts in to caching"""
    import ctypes as C

    import synth
    import torch

    lib = gpu.lib
    lib.spasm_b200_cached_bytes.restype = C.c_longlong
    lib.spasm_b200_set_cache.argtypes = [C.c_int]
    n = 6000
    p, j, x = synth.random_rows(n, n, 10, 42013, 5)
    A = gpu.from_arrays(n, n, p, j, x, 42013)
    free0, _ = torch.cuda.mem_get_info()
    f = gpu.echelonize(A)
    K = gpu.kernel(f)
    assert lib.spasm_b200_cached_bytes() <= (128 << 20), lib.spasm_b200_cached_bytes()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 <= (256 << 20), (free0, free1)
    lib.spasm_b200_set_cache(1)
    try:
        gpu.echelonize(A)
        assert lib.spasm_b200_cached_bytes() > 0
    finally:
        lib.spasm_b200_set_cache(0)
    assert lib.spasm_b200_cached_bytes() <= (128 << 20)
