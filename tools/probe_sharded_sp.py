"""single-process sharded state probe: QFT-(n+g) over 2^g GPUs vs QFT-n on one GPU, wall-clock around run+sync"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C

def run(nq, devices, reps=3):
    sv = cb.StateVector(nq, "double", devices=devices)
    ins = C.qft(nq, C.qft_x(nq, seed=7))
    out = []
    for rep in range(reps):
        sv.sync(); sv.reset()
        a0 = sv.stats(); d0 = sv.dist_stats() if len(devices) > 1 else None
        t0 = time.perf_counter()
        sv.run(ins)
        t1 = time.perf_counter()
        sv.flush()
        t2 = time.perf_counter()
        sv.sync()
        t3 = time.perf_counter()
        a1 = sv.stats(); d1 = sv.dist_stats() if len(devices) > 1 else None
        out.append({"queue_ms": (t1 - t0) * 1e3, "flush_ms": (t2 - t1) * 1e3, "wait_ms": (t3 - t2) * 1e3, "passes": a1["passes"] - a0["passes"],
                    "swap_ms": (d1["swap_ms"] - d0["swap_ms"]) if d1 else 0, "swaps": (d1["swaps"] - d0["swaps"]) if d1 else 0})
    sv.close()
    print(json.dumps({"qubits": nq, "devices": devices, "runs": out}), flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    ng = cb.device_count()
    run(n, [0])
    if ng >= 2:
        run(n, [1])
        run(n + 1, [0, 1])
    if ng >= 4:
        run(n + 2, [0, 1, 2, 3])
    if ng >= 8:
        run(n + 3, list(range(8)))
