"""tcgen05 int8-limb GEMM mod p (dense_mma.cu) against exact integer arithmetic and against the
CUDA-core kernel — bit-exact, ragged shapes included."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(gpu, prime, A, B, Cm, subtract, path):
    M, K = A.shape
    N = B.shape[0]
    f = gpu.lib.spasm_b200_gemm_nt_host
    f.restype = C.c_int
    f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    out = np.ascontiguousarray(Cm, dtype=np.uint32).copy()
    a32, b32 = np.ascontiguousarray(A, dtype=np.uint32), np.ascontiguousarray(B, dtype=np.uint32)  # keep alive across the call
    ms = C.c_double(0)
    used = f(prime, M, N, K, a32.ctypes.data, b32.ctypes.data, out.ctypes.data, int(subtract), path, C.byref(ms))
    assert used >= 0
    return out, used, ms.value


def _exact_product(A, B, prime):
    """(A @ B^T) mod prime, exactly, for residues up to 2^32 (16-bit halves keep every partial sum below 2^63)"""
    M, K = A.shape
    prod = np.zeros((M, B.shape[0]), dtype=np.int64)
    if prime < (1 << 16):
        for k0 in range(0, K, 256):
            prod = (prod + A[:, k0 : k0 + 256] @ B[:, k0 : k0 + 256].T) % prime
        return prod
    Ah, Al = A >> 16, A & 0xFFFF
    for k0 in range(0, K, 256):
        Bk = B[:, k0 : k0 + 256].T
        Bh, Bl = Bk >> 16, Bk & 0xFFFF
        hh = (Ah[:, k0 : k0 + 256] @ Bh) % prime
        hl = (Ah[:, k0 : k0 + 256] @ Bl + Al[:, k0 : k0 + 256] @ Bh) % prime
        ll = (Al[:, k0 : k0 + 256] @ Bl) % prime
        w16 = (1 << 16) % prime
        w32 = (w16 * w16) % prime
        part = ((hh.astype(object) * w32 + hl.astype(object) * w16 + ll.astype(object)) % prime).astype(np.int64)
        prod = (prod + part) % prime
    return prod


# 2 limbs (p < 2^16), 3 limbs (p < 2^24), 4 limbs (up to the largest prime the reference admits, src/SpaSM.jl:74)
@pytest.mark.parametrize("prime", [42013, 65521, 251, 65537, 16777213, 2147483647, 4294967291])
@pytest.mark.parametrize("shape", [(256, 256, 128), (300, 517, 200), (1000, 1024, 1000), (129, 2000, 64), (2048, 384, 999)])
def test_gemm_mod_p(gpu, prime, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N)
    A = rng.integers(0, prime, size=(M, K), dtype=np.int64)
    B = rng.integers(0, prime, size=(N, K), dtype=np.int64)
    C0 = rng.integers(0, prime, size=(M, N), dtype=np.int64)
    # extremes: p-1 everywhere in one row/column stresses the accumulators
    A[0, :] = prime - 1
    B[0, :] = prime - 1
    prod = _exact_product(A, B, prime)
    for subtract in (False, True):
        want = (C0 - prod) % prime if subtract else prod
        got_core, used_core, _ = _gemm(gpu, prime, A, B, C0, subtract, 1)
        assert used_core == 0
        assert np.array_equal(got_core.astype(np.int64), want), "CUDA-core GEMM wrong"
        got, used, _ = _gemm(gpu, prime, A, B, C0, subtract, 0)
        assert used == 1, "the tcgen05 kernel did not run for an eligible shape"
        bad = np.argwhere(got.astype(np.int64) != want)
        assert len(bad) == 0, f"tcgen05 GEMM: {len(bad)} mismatches, first {bad[:4].tolist()} got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}"
