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
    if kind in ("gemm", "c3b", "dense", "dtail"):
        continue
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

def gemm_bench(M, N, K, prime=42013):
    f = g.lib.spasm_b200_gemm_nt_host
    f.restype = C.c_int
    f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    rng = np.random.default_rng(1)
    A = rng.integers(0, prime, size=(M, K), dtype=np.uint32)
    B = rng.integers(0, prime, size=(N, K), dtype=np.uint32)
    Cm = rng.integers(0, prime, size=(M, N), dtype=np.uint32)
    st = (C.c_double * 4)()
    g.lib.spasm_b200_mma_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    for path in (0, 1):
        best = 1e9
        for rep in range(3):
            ms = C.c_double(0)
            g.lib.spasm_b200_mma_stats(st, 1)
            used = f(prime, M, N, K, A.ctypes.data, B.ctypes.data, Cm.ctypes.data, 1, path, C.byref(ms))
            g.lib.spasm_b200_mma_stats(st, 0)
            kms = st[0] if used else ms.value
            best = min(best, kms)
        macs = M * N * K
        print(f"== gemm {M}x{N}x{K} path={'tcgen05' if used else 'cuda-core'}: kernel {best:.3f} ms  {macs/best/1e9:.1f} G modular MAC/s"
              + (f"  = {8*macs/best/1e9:.0f} T int8 OP/s" if used else ""), flush=True)


for arg in sys.argv[1:]:
    if arg.startswith("gemm:"):
        gemm_bench(*map(int, arg[5:].split(",")))

for arg in sys.argv[1:]:
    kind, _, size = arg.partition(":")
    if kind == "c3b":  # banded planted-rank GL7d19-shaped, scale 1/size
        s_ = int(size)
        n, m, r = 1911130 // s_, 1955309 // s_, 1033568 // s_
        t = time.time()
        pjx = synth.banded_planted(n, m, r, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
        print(f"   generated in {time.time()-t:.1f}s nnz/row={pjx[0][-1]/n:.1f} expected rank {r}", flush=True)
        f = run(arg, n, m, *pjx)
        assert f.r == r, (f.r, r)
        L = (C.c_longlong * 7)()
        g.lib.spasm_b200_last_stats.argtypes = [C.POINTER(C.c_longlong)]
        g.lib.spasm_b200_last_stats(L)
        if L[6] > 0:
            print(f"   last Schur: bytes={L[0]/1e9:.3f} GB macs={L[1]/1e9:.3f} G rows={L[2]} light={L[3]} medium={L[4]} heavy={L[5]} "
                  f"time={L[6]/1e3:.1f} ms -> {L[0]/(L[6]*1e-6)/1e9:.1f} GB/s algorithmic", flush=True)
    elif kind == "dense":
        n = int(size)
        D = synth.dense_random(n, n, 65521, 0x5A5A0004)
        import scipy.sparse as sp
        A = g.CSR(sp.csc_matrix(D.T.astype(np.int64)), 65521)
        t = time.time()
        f = g.echelonize(A)
        dt = time.time() - t
        T = (C.c_double * 16)()
        g.lib.spasm_b200_last_timings(T)
        print(f"== dense {n}x{n} mod 65521: rank={f.r} wall={dt:.3f}s  LU-equivalent {2*n**3/3/dt/1e12:.2f} T mod-p OP/s", flush=True)
        print("   " + " ".join(f"{k}={v:.4g}" for k, v in zip(NAMES, T)), flush=True)

for arg in sys.argv[1:]:
    kind, _, size = arg.partition(":")
    if kind == "dtail":  # BASELINE configs[3]: dense n x n mod 65521 through the blocked dense tail
        n = int(size)
        f = g.lib.spasm_b200_dense_tail_bench
        f.restype = C.c_int
        f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
        st = (C.c_double * 4)()
        g.lib.spasm_b200_mma_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
        for rep in range(3):
            ms = C.c_double(0)
            g.lib.spasm_b200_mma_stats(st, 1)
            r = f(65521, n, n, 1000, 0x5A5A0004, C.byref(ms))
            g.lib.spasm_b200_mma_stats(st, 0)
            print(f"== dense tail {n}x{n} mod 65521 (block 1000): rank={r} {ms.value:.1f} ms  -> {2*n**3/3/(ms.value*1e-3)/1e9:.0f} G mod-p OP/s (2n^3/3 / t);"
                  f" tcgen05 kernel {st[0]:.1f} ms for {st[1]:.3g} modular MACs = {8*st[1]/max(st[0],1e-9)/1e9:.0f} T int8 OP/s", flush=True)
