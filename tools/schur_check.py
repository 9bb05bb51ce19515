"""GPU check (not a test): the Schur complement of the 1/SCALE banded instance after one structural round, CUDA library
vs oracle: rows bit for bit and the work counters.  usage: [SPASM_B200_SCHUR_DENSE=0] python tools/schur_check.py [SCALE]"""
import ctypes as C
import sys
import time

sys.path[:0] = [".", "tests"]
import numpy as np

import __graft_entry__ as e
import synth
from test_gpu_engine import structural_round

sc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pkg = e.load_package()
gpu, ora = pkg.SpaSM(), pkg.SpaSM(e.build_oracle())
gpu.log(False), ora.log(False)
ns, ms_, rs = 1911130 // sc, 1955309 // sc, 1033568 // sc
pjx = synth.banded_planted(ns, ms_, rs, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
A = ora.from_arrays(ns, ms_, *pjx, 42013)
lu, U, qinv, p, npiv = structural_round(pkg, ora, A)
rest = np.ascontiguousarray(p[npiv:])
nrest = ns - npiv
pp = rest.ctypes.data_as(C.POINTER(C.c_int32))
res = {}
for name, api in (("oracle", ora), ("gpu", gpu)):
    pout = np.zeros(max(nrest, 1), dtype=np.int32)
    t = time.perf_counter()
    S = pkg.CSR(api, api.lib.spasm_schur(A.data, pp, nrest, C.byref(lu), 0.0, None, None, pout.ctypes.data_as(C.POINTER(C.c_int32))))
    dt = time.perf_counter() - t
    res[name] = (S.arrays(), dt)
ob = C.c_int64.in_dll(ora.lib, "spasm_b200_last_bytes").value
om = C.c_int64.in_dll(ora.lib, "spasm_b200_last_macs").value
Ls = (C.c_longlong * 7)()
gpu.lib.spasm_b200_last_stats.argtypes = [C.POINTER(C.c_longlong)]
gpu.lib.spasm_b200_last_stats(Ls)
same = all(np.array_equal(a, b) for a, b in zip(res["oracle"][0], res["gpu"][0]))
print(f"rows {nrest}, pivots {npiv}; S identical: {same}; oracle {ob} B / {om} MACs in {res['oracle'][1]:.2f}s; "
      f"gpu {Ls[0]} B / {Ls[1]} MACs, smem {Ls[3]} global {Ls[4]} dense {Ls[5]} rows, {Ls[6]/1e3:.1f} ms kernel time "
      f"({Ls[0]/max(Ls[6],1)/1e3:.1f} GB/s algorithmic), call {res['gpu'][1]:.3f}s")
