"""GPU probe: BASELINE configs[3] (dense n x n mod 65521 through the dense tail).  usage: python tools/dense_probe.py [n]"""
import ctypes as C
import sys

sys.path[:0] = [".", "tests"]
import __graft_entry__ as e

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
gpu = e.load_package().SpaSM()
gpu.log(False)
f = gpu.lib.spasm_b200_dense_tail_bench
f.restype = C.c_int
f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
gpu.lib.spasm_b200_set_cache.argtypes = [C.c_int]
gpu.lib.spasm_b200_set_cache(1)
for it in range(3):
    ms = C.c_double(0)
    r = f(65521, n, n, 1000, 0x5A5A0004, C.byref(ms))
    print(f"rank {r} in {ms.value:.1f} ms", flush=True)
