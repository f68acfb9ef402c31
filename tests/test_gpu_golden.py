"""GPU (-m gpu): the engine, through the C ABI, on the committed golden fixtures (wire JSON built by the
reference's own Python front end; expected numbers from the oracle, analytic where available)."""
import json
import os

import numpy as np
import pytest

import cunqa_b200 as cb

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


def sv(rec, key="oracle_statevector"):
    a = np.asarray(rec[key], dtype=float)
    return a[:, 0] + 1j * a[:, 1]


@pytest.fixture(scope="module")
def backend(lib_built):
    be = cb.Backend(0)
    yield be
    be.close()


@pytest.mark.parametrize("name", ["ghz8", "qft7", "vocab5", "hea6"])
@pytest.mark.parametrize("precision,tol", [("double", 1e-10), ("single", 1e-5)])
def test_static_amplitudes(lib_built, name, precision, tol):
    rec = load(name)
    n = rec["task"]["config"]["num_qubits"]
    with cb.StateVector(n, precision) as s:
        s.run(rec["task"]["instructions"])
        assert np.abs(s.amplitudes() - sv(rec)).max() < tol


@pytest.mark.parametrize("name", ["ghz8", "qft7", "vocab5", "hea6", "dyn3"])
def test_counts_equal_oracle(backend, name):
    rec = load(name)
    rec["task"]["config"]["b200_defer_measurement"] = False  # dyn3: the per-shot interpreter is shot-identical
    res = backend.execute(rec["task"])
    assert "ERROR" not in res, res
    if any(i["name"] == "swap" for i in rec["task"]["instructions"]):
        # uncontrolled swaps are pure relabelling in the engine, so its inverse-CDF walks the amplitudes in
        # physical order: same distribution, different shots for the same uniform draws -> shot-noise bound
        p = np.abs(sv(rec)) ** 2
        shots = rec["task"]["config"]["shots"]
        emp = np.zeros_like(p)
        for k, v in res["counts"].items():
            emp[int(k, 2)] = v / shots
        assert sum(res["counts"].values()) == shots
        assert 0.5 * np.abs(emp - p).sum() < 0.5 * np.sqrt(len(p) / shots) + 0.05
    else:
        assert res["counts"] == rec["oracle_counts"]


def test_parameter_update_roundtrip(backend):
    rec = load("hea6")
    backend.execute(rec["task"])
    res = backend.execute(rec["update"])
    assert "ERROR" not in res and sum(res["counts"].values()) == rec["update"]["shots"]


@pytest.mark.parametrize("name", ["teledata", "telegate"])
def test_quantum_communication_families(backend, name):
    rec = load(name)
    from oracle import oracle_np as O
    # default mode (measurements deferred): distribution-equal
    big = [dict(t, config=dict(t["config"], shots=20000)) for t in rec["tasks"]]
    res = backend.execute(big)
    assert "ERROR" not in res, res
    assert backend.last_mode() == 1
    for tid, want in rec["oracle_id_counts"].items():
        assert O.tvd(res["id_counts"][tid], want) < 0.5 * (4 / 400) ** 0.5 + 0.06
    # per-shot interpreter: shot-identical
    for t in rec["tasks"]:
        t["config"]["b200_defer_measurement"] = False
    res = backend.execute(rec["tasks"])
    assert "ERROR" not in res, res
    if name == "teledata":  # qrecv's swap is a relabelling in the engine: distribution-equal, see test_gpu_backend
        for tid, want in rec["oracle_id_counts"].items():
            assert O.tvd(res["id_counts"][tid], want) < 0.5 * (4 / 400) ** 0.5 + 0.06
    else:
        assert res["id_counts"] == rec["oracle_id_counts"]
