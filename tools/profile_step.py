"""One resident echelonization of BASELINE configs[1] (or --rows N) for ncu: no warm-up, no CPU leg.
usage: ncu ... python tools/profile_step.py [--rows N] [--steps K]"""
import argparse
import ctypes as C
import sys

sys.path[:0] = [".", "tests"]
import __graft_entry__ as e
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=bench.FULL_N)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
pkg = e.load_package()
gpu = pkg.SpaSM()
gpu.log(False)
lib = gpu.lib
lib.spasm_b200_upload.restype = C.c_void_p
lib.spasm_b200_upload.argtypes = [C.c_void_p]
lib.spasm_b200_echelonize_resident.restype = C.c_int
lib.spasm_b200_echelonize_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
lib.spasm_b200_set_cache.argtypes = [C.c_int]
lib.spasm_b200_set_cache(1)
p, j, x = bench.make_input(a.rows)
A = gpu.from_arrays(a.rows, a.rows, p, j, x, bench.PRIME)
h = lib.spasm_b200_upload(C.cast(A.data, C.c_void_p))
opts = gpu.EchelonizeOpts()
for _ in range(a.steps):
    ms = C.c_double(0)
    r = lib.spasm_b200_echelonize_resident(h, C.byref(opts), C.byref(ms))
    print(f"rank {r} in {ms.value / 1e3:.3f} s", flush=True)
