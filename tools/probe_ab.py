"""A/B probe of library builds on one box: CB200_PKG_ROOT=<dir holding a cunqa_b200 package> python tools/probe_ab.py <circuits> [n]
(default: the repo's own package).  Wall clock around run + flush + sync after a lazy reset (the |0> initialisation
kernel, 1*S, is inside the timed region for every build alike), best of 3."""
import sys, os, time, json
root = os.environ.get("CB200_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.abspath(root))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C


def probe(name, ins, n, precision="double", reps=3):
    sv = cb.StateVector(n, precision)
    best = None
    for r in range(reps):
        sv.reset()
        s0 = sv.stats()
        t0 = time.perf_counter()
        sv.run(ins)
        sv.flush()
        sv.sync()
        dt = time.perf_counter() - t0
        passes = sv.stats()["passes"] - s0["passes"]
        best = dt if best is None else min(best, dt)
    sv.close()
    print(json.dumps({"pkg": os.path.relpath(root), "probe": name, "n": n, "precision": precision, "gates": C.n_gates(ins), "passes": passes,
                      "ms": round(best * 1e3, 2), "gates_per_s": round(C.n_gates(ins) / best, 1)}), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rng = np.random.default_rng(1)
    for which in sys.argv[1].split(","):
        if which == "qv": probe("qv_full", C.quantum_volume(n), n)
        elif which == "qv32": probe("qv_full_fp32", C.quantum_volume(n), n, "single")
        elif which == "qft": probe("qft", C.qft(n, C.qft_x(n)), n)
        elif which == "hea": probe("hea4", C.hea_ansatz(n, 4, measure=False), n)
        elif which == "local3":
            a, b, c4, d = n - 1, n - 2, n - 3, n - 4
            probe("local3", [C.unitary_inst(C.haar_unitary(rng, 4), list(pq)) for pq in [(a, b), (c4, d), (a, c4), (b, d), (a, d), (b, c4)]], n)
        elif which == "ghz": probe("ghz", C.ghz(n, measure=False), n)
