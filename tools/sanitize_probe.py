"""small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): static + diagonal runs +
look-ahead scheduling at 20 qubits, deferred measurement with qubit recycling, the per-shot interpreter, sampling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O
from test_oracle import _many_remote_ops, _nested_and_multibit_conditions

be = cb.Backend(0)
n = 20
for prec in ("double", "single"):
    r = be.execute(C.task(C.qft(n, C.qft_x(n)) + C.measure_all(n), n, shots=64, seed=1, precision=prec))
    assert "ERROR" not in r, r
    r = be.execute(C.task(C.quantum_volume(n, depth=3, seed=2) + C.measure_all(n), n, shots=64, seed=1, precision=prec))
    assert "ERROR" not in r, r
r = be.execute(C.task(C.hea_ansatz(12, 3, 1, measure=True), 12, shots=128, seed=3))
assert "ERROR" not in r, r
tasks = _many_remote_ops(2)
msg = [C.task(t["instructions"], t["num_qubits"], shots=256, seed=5, task_id=t["id"], is_dynamic=True,
              b200_teledata_corrections="protocol") for t in tasks]
r = be.execute(msg); assert "ERROR" not in r and be.last_mode() == 1, r
for m in msg: m["config"]["b200_defer_measurement"] = False
r = be.execute(msg); assert "ERROR" not in r and be.last_mode() == 2, r
t = _nested_and_multibit_conditions()[0]
r = be.execute(C.task(t["instructions"], 4, num_clbits=4, shots=256, seed=4, is_dynamic=True)); assert "ERROR" not in r, r
with cb.StateVector(14) as sv:
    sv.run(C.quantum_volume(14, depth=4, seed=1))
    sv.probabilities([0, 3, 5]); sv.prob1(2); sv.norm(); sv.sample(100, 7); sv.amplitudes(np.arange(64, dtype=np.uint64))
# round 2: independent registers batched in shared launches (per-register matrices and phase-fan tables), a run of
# registers with different structures, the staged diagonal sub-tables
res = be.execute_many(C.config5_groups(6, n_each=5, layers=2, shots=64, seed=3))
assert all("ERROR" not in x for x in res), res
res = be.execute_many([C.task(C.qft(13, 77) + C.measure_all(13), 13, shots=32, seed=g) for g in range(3)] +
                      [C.task(C.ghz(13), 13, shots=32, seed=9)])
assert all("ERROR" not in x for x in res) and be.batch_info()["uniform_flushes"] >= 1, (res, be.batch_info())
os.environ["CB200_STAGE_DIAG"] = "1"
with cb.StateVector(16) as sv:
    x = C.qft_x(16)
    sv.run(C.qft(16, x))
    idx = np.arange(1 << 16, dtype=np.uint64)
    assert np.abs(sv.amplitudes(idx) - O.qft_amplitudes(16, x, idx)).max() < 1e-12
del os.environ["CB200_STAGE_DIAG"]
be.close()
print("sanitize probe ok")
