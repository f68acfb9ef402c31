"""config 1: GHZ-20, 1024 shots, seed 1234, end to end through cb200_backend_execute (latency-bound)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunqa_b200 as cb
from cunqa_b200 import circuits as C
n, shots = 20, 1024
msg = json.dumps(C.task(C.ghz(n), n, shots=shots, seed=1234))
be = cb.Backend(0)
be.execute(msg)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); res = be.execute(msg); ts.append(time.perf_counter() - t0)
ts.sort()
print(json.dumps({"probe": "ghz20", "shots": shots, "median_ms": round(ts[len(ts) // 2] * 1e3, 3), "min_ms": round(ts[0] * 1e3, 3),
                  "shots_per_s": round(shots / ts[len(ts) // 2], 1), "circuits_per_s": round(1 / ts[len(ts) // 2], 1),
                  "counts": res["counts"], "stats": be.stats()}))
