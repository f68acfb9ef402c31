"""A/B of library builds: python tools/probe_variant.py <path/to/libcunqa_b200.so> [n] -- QV-n (full depth) through the
state-level C ABI (only entry points every build has), wall clock around flush + sync, best of 3."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cunqa_b200 import circuits as C

lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
vp = ctypes.c_void_p
lib.cb200_state_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
lib.cb200_apply_matrix.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
for f in ("cb200_flush", "cb200_sync", "cb200_state_reset", "cb200_state_destroy"):
    getattr(lib, f).argtypes = [vp]
lib.cb200_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
lib.cb200_last_error.restype = ctypes.c_char_p
h = vp()
assert lib.cb200_state_create(ctypes.byref(h), n, 64, 0, 1) == 0, lib.cb200_last_error()
ins = C.quantum_volume(n, seed=2024)
ops = []
for i in ins:
    a = np.asarray(i["matrix"][0], dtype=float)
    m = np.ascontiguousarray(a[..., 0] + 1j * a[..., 1])
    ops.append(((ctypes.c_int * 2)(*i["qubits"]), m))
best = None
for rep in range(4):
    lib.cb200_state_reset(h)
    for q, m in ops:
        assert lib.cb200_apply_matrix(h, q, 2, None, 0, m.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    t0 = time.perf_counter()
    assert lib.cb200_sync(h) == 0, lib.cb200_last_error()
    dt = time.perf_counter() - t0
    if rep:
        best = dt if best is None else min(best, dt)
st = (ctypes.c_uint64 * 8)()
lib.cb200_stats(h, st)
print(json.dumps({"lib": sys.argv[1], "n": n, "gates": len(ops), "ms": round(best * 1e3, 2), "gates_per_s": round(len(ops) / best, 1), "passes_total": int(st[0])}), flush=True)
lib.cb200_state_destroy(h)
