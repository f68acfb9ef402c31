"""quick device-timed probe of the tile-pass kernel (not the bench): ms per pass and GB/s."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C

def probe(name, ins, n, precision="double", reps=2):
    sv = cb.StateVector(n, precision)
    bytes_amp = 16 if precision == "double" else 8
    out = []
    for r in range(reps):
        sv.reset()
        s0 = sv.stats()
        t0 = time.perf_counter()
        sv.run(ins)
        t_queue = time.perf_counter() - t0
        sv.flush()
        t_flush = time.perf_counter() - t0 - t_queue
        sv.sync()
        dt = time.perf_counter() - t0
        s1 = sv.stats()
        passes = s1["passes"] - s0["passes"]
        S = bytes_amp * (1 << n)
        out.append((dt, passes, 2 * S * passes / dt / 1e9))
    sv.close()
    dt, passes, gbs = out[-1]
    print(json.dumps({"probe": name, "n": n, "precision": precision, "gates": C.n_gates(ins), "passes": passes,
                      "ms": round(dt * 1e3, 2), "ms_per_pass": round(dt * 1e3 / max(passes, 1), 3),
                      "GBps": round(gbs, 1), "gates_per_s": round(C.n_gates(ins) / dt, 1), "host_queue_ms": round(t_queue * 1e3, 1), "host_flush_ms": round(t_flush * 1e3, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("CB200_")}}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    if which in ("qv", "all"):
        probe("qv", C.quantum_volume(n, depth=min(n, 10)), n)
    if which == "qvfull":
        probe("qv_full", C.quantum_volume(n), n)
    if which in ("qft", "all"):
        probe("qft", C.qft(n, C.qft_x(n)), n)
    if which in ("h", "all"):
        probe("h_layer", [{"name": "h", "qubits": [q]} for q in range(n)], n)
    if which == "local3":  # fixed pass: 3 register-block pairs on 4 high qubits (isolates kernel efficiency)
        import numpy as np
        rng = np.random.default_rng(1)
        a, b, c4, d = n - 1, n - 2, n - 3, n - 4
        pairs = [(a, b), (c4, d), (a, c4), (b, d), (a, d), (b, c4)]
        reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        probe("local3", [C.unitary_inst(C.haar_unitary(rng, 4), list(pq)) for _ in range(reps) for pq in pairs], n, reps=3)
    if which == "x1":   # streaming ceiling of the tile pipeline: one flop-free op per pass
        probe("x1", [{"name": "x", "qubits": [n - 1]}], n, reps=3)
    if which == "rz1":
        probe("rz1", [{"name": "rz", "qubits": [n - 1], "params": [0.3]}], n, reps=3)
    if which == "u2q":  # a single 2-qubit dense gate per pass
        import numpy as np
        probe("u2q", [C.unitary_inst(C.haar_unitary(np.random.default_rng(0), 4), [n - 1, n - 2])], n, reps=3)
    if which in ("qv32", "all"):
        probe("qv_fp32", C.quantum_volume(n, depth=min(n, 10)), n, "single")
