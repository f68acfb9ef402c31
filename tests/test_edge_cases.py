"""Edge cases of the wire format, on the host-only entry points (CPU) and through the engine (GPU)."""
import json

import numpy as np
import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from cunqa_b200.binding import B200Error
from oracle import oracle_np as O


# ---------------------------------------------------------------- host only (no GPU)
def test_plan_of_empty_circuit(lib_built):
    plan = cb.plan(C.task([], 3))
    assert plan["ops"] == [] and plan["passes"] == [] and plan["gates"] == 0


def test_plan_only_measurements_and_barriers(lib_built):
    ins = [{"name": "barrier", "qubits": [0, 1]}, {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "id", "qubits": [1]}]
    assert cb.plan(C.task(ins, 2))["ops"] == []


@pytest.mark.parametrize("bad,msg", [
    ('{"id":"x","config":{"num_qubits":2},"instructions":[{"name":"h"}]}', "needs at least"),
    ('{"id":"x","config":{"num_qubits":2},"instructions":[{"qubits":[0]}]}', "name"),
    ('{"id":"x","instructions":[]}', "config"),
    ('{"id":"x","config":{"num_qubits":0},"instructions":[]}', "num_qubits"),
    ('{"id":"x","config":{"num_qubits":2},"instructions":[{"name":"unitary","qubits":[0],"matrix":[[[[1,0],[0,0]],[[0,0],[1,0]],[[0,0],[0,0]]]]}]}', "square|power of two|match"),
    ('{"id":"x","config":{"num_qubits":2},"instructions":[{"name":"rx","qubits":[0],"params":[]}]}', "parameter"),
    ('not json at all', "JSON|unexpected"),
])
def test_malformed_tasks_are_rejected(lib_built, bad, msg):
    import re
    with pytest.raises(B200Error) as e:
        cb.plan(bad)
    assert re.search(msg, str(e.value)), str(e.value)


def test_sixty_qubit_plan_is_host_only(lib_built):
    """planning is independent of the state size (the scheduler works on qubit masks)"""
    n = 60
    plan = cb.plan(C.task(C.ghz(n, measure=False), n))
    assert len(plan["ops"]) <= n and len(plan["passes"]) >= 1


def test_wide_controls_do_not_cost_tile_slots(lib_built):
    n = 20
    ins = [{"name": "mcx", "qubits": list(range(15)) + [19]}]  # 15 controls
    plan = cb.plan(C.task(ins, n), {"tile_bits": 8, "min_run_bits": 3})
    assert len(plan["passes"]) == 1 and plan["ops"][0]["kind"] == "permx" and len(plan["ops"][0]["ctrl"]) == 15


# ---------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(lib_built):
    be = cb.Backend(0)
    yield be
    be.close()


@gpu
def test_empty_circuit_counts(backend):
    res = backend.execute(C.task([], 3, shots=17, seed=1))
    assert res["counts"] == {"000": 17}


@gpu
def test_zero_shots(backend):
    res = backend.execute(C.task(C.ghz(3), 3, shots=0, seed=1))
    assert res["counts"] == {} and "ERROR" not in res


@gpu
def test_ragged_measurement_maps(backend):
    # one qubit into two clbits, an unmeasured clbit, more clbits than qubits, a qubit measured twice
    ins = [{"name": "x", "qubits": [1]}, {"name": "h", "qubits": [0]},
           {"name": "measure", "qubits": [1], "clbits": [0]}, {"name": "measure", "qubits": [1], "clbits": [4]},
           {"name": "measure", "qubits": [0], "clbits": [2]}, {"name": "measure", "qubits": [0], "clbits": [3]}]
    res = backend.execute(C.task(ins, 2, num_clbits=6, shots=200, seed=3))
    _, want = O.run_static(ins, 2, n_clbits=6, shots=200, seed=3, backend="numpy")
    assert res["counts"] == want
    assert set(res["counts"]) <= {"010001", "011101"}  # clbits 2,3 agree; clbits 0,4 = 1; clbits 1,5 never written


@gpu
def test_errors_inside_execute_become_error_json(backend):
    for ins, frag in [([{"name": "h", "qubits": [7]}], "out of range"), ([{"name": "cx", "qubits": [0, 0]}], "duplicate"),
                      ([{"name": "fs", "qubits": [0, 1], "params": [0.1, 0.2]}], "unsupported"),
                      ([{"name": "measure", "qubits": [0], "clbits": [9]}], "clbit")]:
        res = backend.execute(C.task(ins, 2, shots=5))
        assert "ERROR" in res and frag in res["ERROR"], res
    # the backend survives and keeps working
    assert backend.execute(C.task(C.ghz(2), 2, shots=4, seed=1))["counts"].keys() <= {"00", "11"}


@gpu
def test_register_too_large_is_an_error_not_a_crash(backend):
    res = backend.execute(C.task([{"name": "h", "qubits": [0]}], 38, shots=1))
    assert "ERROR" in res


@gpu
def test_million_shots(backend):
    n = 6
    ins = C.hea_ansatz(n, 2)
    res = backend.execute(C.task(ins, n, shots=1_000_000, seed=8))
    st, _ = O.run_static(ins, n, backend="numpy")
    p = np.abs(st.amplitudes()) ** 2
    emp = np.zeros(1 << n)
    for k, v in res["counts"].items():
        emp[int(k, 2)] = v / 1e6
    assert 0.5 * np.abs(emp - p).sum() < 0.5 * np.sqrt((1 << n) / 1e6) + 2e-3


@gpu
def test_single_qubit_register(backend):
    res = backend.execute(C.task([{"name": "x", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]}], 1, shots=9, seed=1))
    assert res["counts"] == {"1": 9}


@gpu
def test_dynamic_edge_cases(backend):
    # cif on a never-written clbit (reads 0), conditioned measurement, copy
    ins = [{"name": "h", "qubits": [0]},
           {"name": "cif", "clbits": [1], "operation": "and", "condition": 0, "instructions": [{"name": "measure", "qubits": [0], "clbits": [0]}]},
           {"name": "copy", "l_clbits": [2], "r_clbits": [0]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [1]}]},
           {"name": "measure", "qubits": [1], "clbits": [1]}]
    res = backend.execute(C.task(ins, 2, num_clbits=3, shots=300, seed=12, is_dynamic=True, b200_defer_measurement=False))
    want = O.run_dynamic([{"id": "task", "num_qubits": 2, "num_clbits": 3, "instructions": ins}], 300, seed=12)["task"]
    assert res["counts"] == want and set(want) <= {"000", "111"}
    dfr = backend.execute(C.task(ins, 2, num_clbits=3, shots=3000, seed=12, is_dynamic=True))  # measurements deferred
    assert backend.last_mode() == 1 and set(dfr["counts"]) == {"000", "111"} and abs(dfr["counts"]["111"] / 3000 - 0.5) < 0.05
