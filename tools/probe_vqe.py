"""parameter-update loop probe (the `upgrade_parameters` path, cunqa/qjob.py:183-236): one HEA task, then a stream of
{"params": [...]} messages; wall time per execute vs the engine's own time_taken."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    ins = C.hea_ansatz(n, layers, 1, measure=True)
    npar = sum(len(i.get("params", [])) for i in ins)
    be = cb.Backend(0)
    t0 = time.perf_counter()
    res = be.execute(C.task(ins, n, shots=1024, seed=1))
    first = time.perf_counter() - t0
    assert "ERROR" not in res, res
    rng = np.random.default_rng(0)
    walls, inner = [], []
    for it in range(iters):
        msg = json.dumps({"params": list(rng.uniform(0, 6.28, npar)), "shots": 1024})
        t0 = time.perf_counter()
        res = be.execute(msg)
        walls.append(time.perf_counter() - t0)
        inner.append(res["time_taken"])
        assert "ERROR" not in res, res
    # the C ABI alone: message bytes in, result text out (no Python-side JSON work inside the timed region)
    msgs = [json.dumps({"params": list(rng.uniform(0, 6.28, npar)), "shots": 1024}).encode() for _ in range(iters)]
    t0 = time.perf_counter()
    for m in msgs:
        be.execute(m, parse=False)
    abi = (time.perf_counter() - t0) / iters
    st = be.stats()
    print(json.dumps({"probe": "vqe_loop", "c_abi_ms_per_eval": round(abi * 1e3, 3), "c_abi_evals_per_s": round(1 / abi, 1), "n": n, "layers": layers, "gates": C.n_gates(ins), "params": npar, "first_ms": round(first * 1e3, 3),
                      "wall_ms_median": round(float(np.median(walls)) * 1e3, 3), "engine_ms_median": round(float(np.median(inner)) * 1e3, 3),
                      "passes": st["passes"], "launches": st["launches"], "evals_per_s": round(1 / float(np.median(walls)), 1)}), flush=True)
    be.close()

if __name__ == "__main__":
    main()
