"""GPU-side probe: time echelonize (+kernel) at several scales and print the phase breakdown."""
import ctypes as C
import sys
import time

sys.path[:0] = [".", "tests"]
import numpy as np

import __graft_entry__ as e
import synth

pk = e.load_package()
g = pk.SpaSM()
g.lib.spasm_b200_last_timings.argtypes = [C.POINTER(C.c_double)]
NAMES = "total upload FL FLcol greedy reorder+extract density schur tail download rounds flcol_rounds greedy_windows schur_bytes schur_macs schur_ms".split()


def run(tag, n, m, p, j, x, prime=42013, kernel=False, **kw):
    A = g.from_arrays(n, m, p, j, x, prime)
    t = time.time()
    f = g.echelonize(A, **kw)
    dt = time.time() - t
    T = (C.c_double * 16)()
    g.lib.spasm_b200_last_timings(T)
    print(f"== {tag}: {n}x{m} nnz={p[-1]} rank={f.r} nnzU={f.U.nnz()} wall={dt:.3f}s", flush=True)
    print("   " + " ".join(f"{k}={v:.4g}" for k, v in zip(NAMES, T)), flush=True)
    if kernel:
        t = time.time()
        K = g.kernel(f)
        print(f"   kernel: {K.shape} nnz={K.nnz()} wall={time.time()-t:.3f}s", flush=True)
    return f


for arg in sys.argv[1:]:
    kind, _, size = arg.partition(":")
    s = int(size)
    if kind == "c1":
        run(arg, s, s, *synth.random_rows(s, s, 5, 42013, 0x5A5A0001), kernel=True)
    elif kind == "c2":
        run(arg, s, s, *synth.random_rows(s, s, 10, 42013, 0x5A5A0002))
    elif kind == "c2sparse":
        run(arg, s, s, *synth.random_rows(s, s, 10, 42013, 0x5A5A0002), enable_dense=False, max_round=1)
    elif kind == "c3":
        n, m, r = 1911130 // s, 1955309 // s, 1033568 // s
        run(arg, n, m, *synth.planted_rank(n, m, r, 18.0, 42013, 0x5A5A0003))
    elif kind == "c5":
        n, m = 500000 // s, 1000000 // s
        run(arg, n, m, *synth.random_rows(n, m, 8, 42013, 0x5A5A0005), kernel=(s >= 50))
