"""e2e (host structs through spasm_echelonize) against resident timing on a sample of configs[1], one / two streams."""
import ctypes as C
import os
import sys
import time

sys.path[:0] = [".", "tests"]
import __graft_entry__ as e
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pkg = e.load_package()
gpu = pkg.SpaSM()
gpu.log(False)
lib = gpu.lib
lib.spasm_b200_set_cache.argtypes = [C.c_int]
lib.spasm_b200_set_cache(1)
lib.spasm_b200_last_timings.argtypes = [C.POINTER(C.c_double)]
p, j, x = bench.make_input(n)
A = gpu.from_arrays(n, n, p, j, x, bench.PRIME)
for mode in ("two", "one", "two"):
    if mode == "one":
        os.environ["SPASM_B200_ONE_STREAM"] = "1"
    else:
        os.environ.pop("SPASM_B200_ONE_STREAM", None)
    for it in range(3):
        t = time.perf_counter()
        f = gpu.echelonize(A)
        r = f.r
        t1 = time.perf_counter()
        T = (C.c_double * 16)()
        lib.spasm_b200_last_timings(T)
        del f
        t2 = time.perf_counter()
        print(f"{mode} streams, run {it}: rank {r} call {t1 - t:.3f}s free {t2 - t1:.3f}s | lib total {T[0]:.3f} greedy {T[4]:.3f} tail {T[8]:.3f} download {T[9]:.3f}", flush=True)
