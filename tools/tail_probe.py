"""GPU probe of the dense tail alone (spasm_b200_dense_tail_bench: random dense n x n mod p generated on the device):
one stream against two streams, and how many SMs the second stream's tensor-core launches may use.
  python tools/tail_probe.py [n] [prime]"""
import ctypes as C
import os
import sys

sys.path[:0] = [".", "tests"]
import __graft_entry__ as e

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prime = int(sys.argv[2]) if len(sys.argv) > 2 else 42013
pk = e.load_package()
g = pk.SpaSM()
g.log(False)
f = g.lib.spasm_b200_dense_tail_bench
f.restype = C.c_int
f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
g.lib.spasm_b200_set_cache.argtypes = [C.c_int]
g.lib.spasm_b200_set_cache(1)
os.environ["SPASM_B200_PROFILE"] = "2"
configs = [("one stream", {"SPASM_B200_ONE_STREAM": "1"}), ("two streams, 116 SMs", {"SPASM_B200_TWO_STREAMS": "1", "SPASM_B200_AUX_SMS": "116"}),
           ("two streams, 132 SMs", {"SPASM_B200_AUX_SMS": "132"}), ("two streams, 100 SMs", {"SPASM_B200_AUX_SMS": "100"}),
           ("two streams, 148 SMs", {"SPASM_B200_AUX_SMS": "148"})]
for name, env in configs:
    for k in ("SPASM_B200_ONE_STREAM", "SPASM_B200_AUX_SMS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for it in range(2):
        ms = C.c_double(0)
        sys.stderr.flush()
        r = f(prime, n, n, 1000, 1234, C.byref(ms))
        print(f"{name}: run {it}: rank {r}, {ms.value:.1f} ms", flush=True)
