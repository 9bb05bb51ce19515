"""GPU probe (not a test): phase breakdown of the sparse regime on the GL7d19-shaped banded generator.
usage: python tools/sparse_probe.py SCALE [verbose]"""
import ctypes as C
import sys
import time

sys.path[:0] = [".", "tests"]
import __graft_entry__ as e
import bench
import synth

sc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pkg = e.load_package()
gpu = pkg.SpaSM()
gpu.log(len(sys.argv) > 2)
lib = gpu.lib
ns, ms_, rs = 1911130 // sc, 1955309 // sc, 1033568 // sc
t = time.time()
pjx = synth.banded_planted(ns, ms_, rs, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
print(f"generated {ns}x{ms_} nnz {pjx[0][-1]} in {time.time()-t:.1f}s", flush=True)
A = gpu.from_arrays(ns, ms_, *pjx, 42013)
lib.spasm_b200_set_cache.argtypes = [C.c_int]
lib.spasm_b200_set_cache(1)
for it in range(2):
    t = time.perf_counter()
    f = gpu.echelonize(A)
    dt = time.perf_counter() - t
    assert f.r == rs, (f.r, rs)
    T = (C.c_double * 16)()
    lib.spasm_b200_last_timings.argtypes = [C.POINTER(C.c_double)]
    lib.spasm_b200_last_timings(T)
    ph = {k: round(v, 4) for k, v in zip(bench.TIMING_NAMES, T) if v}
    Ls = (C.c_longlong * 7)()
    lib.spasm_b200_last_stats.argtypes = [C.POINTER(C.c_longlong)]
    lib.spasm_b200_last_stats(Ls)
    print(f"run {it}: {dt:.3f}s rank {f.r} phases {ph}")
    print(f"   last schur: bytes {Ls[0]} macs {Ls[1]} rows {Ls[2]} light {Ls[3]} medium {Ls[4]} heavy {Ls[5]} ms {Ls[6]/1000:.2f} -> {Ls[0]/max(Ls[6],1)/1e3:.1f} GB/s", flush=True)
    del f
