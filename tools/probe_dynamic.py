"""config 5 probe: one quantum-communication group = two 11-qubit vQPUs + 2 communication qubits (24-qubit joint
register), hardware-efficient ansatz + teledata / telegate, mid-circuit measurement and feed-forward; shots batched."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunqa_b200 as cb
from cunqa_b200 import circuits as C

def run(kind, n_each, shots, layers=2, **cfg):
    tasks = (C.hea_teledata_pair if kind == "teledata" else C.hea_telegate_pair)(n_each, layers)
    msg = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=99, task_id=t["id"], is_dynamic=True, **cfg) for t in tasks]
    be = cb.Backend(0)
    be.execute(msg)  # warm-up (allocation)
    t0 = time.perf_counter()
    res = be.execute(msg)
    dt = time.perf_counter() - t0
    st = be.stats()
    assert "ERROR" not in res, res
    print(json.dumps({"probe": kind, "joint_qubits": 2 * n_each + 2, "shots": shots, "s": round(dt, 3), "shots_per_s": round(shots / dt, 1),
                      "circuits_per_s": round(1 / dt, 3), "mode": be.last_mode(), "passes": st["passes"], "launches": st["launches"], "cfg": cfg}), flush=True)
    be.close()

if __name__ == "__main__":
    n_each = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    shots = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    run("teledata", n_each, shots)
    run("telegate", n_each, shots)
    if "--both" in sys.argv:  # the batched per-shot interpreter, for comparison
        run("teledata", n_each, shots, b200_defer_measurement=False)
        run("telegate", n_each, shots, b200_defer_measurement=False)
