"""GPU probe: one tcgen05 limb GEMM C -= A.B^T mod p at the depth of the deferred trailing updates.
usage: python tools/gemm_probe.py [M N K prime]"""
import ctypes as C
import sys

sys.path[:0] = [".", "tests"]
import numpy as np

import __graft_entry__ as e

M, N, K, prime = (int(v) for v in (sys.argv[1:5] + ["32768", "16384", "4096", "42013"][len(sys.argv) - 1:]))
gpu = e.load_package().SpaSM()
f = gpu.lib.spasm_b200_gemm_nt_host
f.restype = C.c_int
f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
rng = np.random.default_rng(1)
A = rng.integers(0, prime, size=(M, K), dtype=np.uint32)
B = rng.integers(0, prime, size=(N, K), dtype=np.uint32)
Cm = rng.integers(0, prime, size=(M, N), dtype=np.uint32)
for it in range(2):
    ms = C.c_double(0)
    used = f(prime, M, N, K, A.ctypes.data, B.ctypes.data, Cm.ctypes.data, 1, 0, C.byref(ms))
    L = 2 if prime < 65536 else 3 if prime < (1 << 24) else 4
    print(f"tcgen05={used} {M}x{N}x{K} mod {prime}: {ms.value:.3f} ms incl. limb split = {2.0 * L * L * M * N * K / (ms.value * 1e-3) / 1e12:.0f} T int8 OP/s", flush=True)
