"""hardware-efficient-ansatz probe: n qubits x L layers (ry, rz on every qubit + cx ladder), static evolution"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunqa_b200 as cb
from cunqa_b200 import circuits as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ins = C.hea_ansatz(n, layers, 1, measure=False)
with cb.StateVector(n) as sv:
    best = None
    for rep in range(3):
        sv.reset(); sv.sync()
        s0 = sv.stats()
        t0 = time.perf_counter()
        sv.run(ins); sv.flush(); sv.sync()
        dt = time.perf_counter() - t0
        s1 = sv.stats()
        best = dt if best is None else min(best, dt)
    print(json.dumps({"probe": "hea", "n": n, "layers": layers, "gates": C.n_gates(ins), "passes": s1["passes"] - s0["passes"],
                      "ms": round(best * 1e3, 2), "gates_per_s": round(C.n_gates(ins) / best, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("CB200_")}}), flush=True)
