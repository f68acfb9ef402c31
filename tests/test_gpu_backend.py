"""GPU (-m gpu): whole-task execution through cb200_backend_execute -- the call the CUNQA plugin
makes -- against the oracle's interpreter (static sample path and per-shot dynamic path)."""
import math

import numpy as np
import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(lib_built):
    be = cb.Backend(0)
    yield be
    be.close()


def test_ghz20_config1(backend):
    """BASELINE config 1: GHZ-20, 1024 shots, seed 1234"""
    n, shots = 20, 1024
    res = backend.execute(C.task(C.ghz(n), n, shots=shots, seed=1234))
    assert "ERROR" not in res, res
    counts = res["counts"]
    assert set(counts) <= {"0" * n, "1" * n} and sum(counts.values()) == shots
    assert abs(counts["0" * n] / shots - 0.5) <= 4 * math.sqrt(0.25 / shots)
    assert res["time_taken"] > 0
    _, want = O.run_static(C.ghz(n), n, shots=shots, seed=1234, backend="c")
    assert counts == want  # shared RNG stream


def test_static_counts_equal_oracle(backend):
    n = 9
    ins = C.quantum_volume(n, depth=5, seed=4)
    # measure a subset into shuffled clbits: clbit c is character ncl-1-c
    ins += [{"name": "measure", "qubits": [q], "clbits": [c]} for q, c in [(0, 3), (4, 0), (8, 1)]]
    res = backend.execute(C.task(ins, n, num_clbits=5, shots=3000, seed=21))
    _, want = O.run_static(ins, n, n_clbits=5, shots=3000, seed=21, backend="c")
    assert res["counts"] == want
    assert all(len(k) == 5 for k in res["counts"])


def test_precision_single(backend):
    n = 12
    ins = C.hea_ansatz(n, 3)
    res = backend.execute(C.task(ins, n, shots=4000, seed=3, precision="single"))
    _, want = O.run_static(ins, n, shots=4000, seed=3, backend="c")
    assert O.tvd(res["counts"], want) < 0.01  # rounding can move a handful of shots


def test_save_state(backend):
    n = 5
    ins = C.quantum_volume(n, depth=3, seed=2) + [{"name": "save_state", "qubits": list(range(n))}]
    res = backend.execute(C.task(ins, n, shots=10, seed=1))
    sv = np.array([complex(*z) for z in res["statevector"]])
    ref, _ = O.run_static(ins, n, backend="numpy")
    assert np.abs(sv - ref.amplitudes()).max() < 1e-12


def test_errors_become_error_json(backend):
    res = backend.execute(C.task([{"name": "hz2", "qubits": [0]}], 2))
    assert "ERROR" in res and "unsupported" in res["ERROR"]
    res = backend.execute({"params": [0.1], "shots": 5}) if False else res
    res = backend.execute(C.task([{"name": "h", "qubits": [0]}], 2, method="stabilizer"))
    assert "ERROR" in res and "statevector only" in res["ERROR"]
    res = backend.execute("{not json")
    assert "ERROR" in res


def test_upgrade_parameters_path(lib_built):
    """{"params":[...],"shots":N} re-runs the cached circuit with new angles (quantum_task.cpp:39-108)"""
    be = cb.Backend(0)
    n = 6
    p0 = np.linspace(0.1, 2.0, 2 * n * 2)
    p1 = p0[::-1].copy()
    res = be.execute({"params": list(p1), "shots": 10})
    assert "ERROR" in res and "Circuit not sent" in res["ERROR"]
    be.execute(C.task(C.hea_ansatz(n, 2, params=p0), n, shots=100, seed=5))
    res = be.execute({"params": list(p1), "shots": 2500})
    _, want = O.run_static(C.hea_ansatz(n, 2, params=p1), n, shots=2500, seed=5, backend="numpy")
    assert res["counts"] == want
    be.close()


def dyn_task(ins, n, ncl, shots, seed, **cfg):
    """dynamic task on the batched per-shot interpreter (shot-identical to the oracle); pass
    b200_defer_measurement=True for the default deferred-measurement mode (distribution-identical)"""
    cfg.setdefault("b200_defer_measurement", False)
    return C.task(ins, n, num_clbits=ncl, shots=shots, seed=seed, is_dynamic=True, **cfg)


def test_mid_circuit_measure_and_cif(backend):
    ins = [{"name": "h", "qubits": [0]}, {"name": "ry", "qubits": [2], "params": [0.9]},
           {"name": "measure", "qubits": [0], "clbits": [0]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 1,
            "instructions": [{"name": "x", "qubits": [1]}, {"name": "rx", "qubits": [2], "params": [1.3]}]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 0,
            "instructions": [{"name": "h", "qubits": [1]}]},
           {"name": "measure", "qubits": [1], "clbits": [1]}, {"name": "measure", "qubits": [2], "clbits": [2]},
           {"name": "reset", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [3]}]
    shots = 600
    res = backend.execute(dyn_task(ins, 3, 4, shots, 17))
    assert "ERROR" not in res, res
    want = O.run_dynamic([{"id": "task", "num_qubits": 3, "num_clbits": 4, "instructions": ins}], shots, seed=17)["task"]
    assert res["counts"] == want  # same per-shot RNG counters => identical shots
    assert all(k[0] == "0" for k in res["counts"])  # clbit 3 (leftmost) holds the reset qubit


@pytest.mark.parametrize("op,cond", [("and", 1), ("or", 1), ("xor", 1), ("andn", 0), ("orn", 1), ("xorn", 0)])
def test_cif_operations(backend, op, cond):
    ins = [{"name": "h", "qubits": [0]}, {"name": "h", "qubits": [1]},
           {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "measure", "qubits": [1], "clbits": [1]},
           {"name": "cif", "clbits": [0, 1], "operation": op, "condition": cond, "instructions": [{"name": "x", "qubits": [2]}]},
           {"name": "measure", "qubits": [2], "clbits": [2]}]
    res = backend.execute(dyn_task(ins, 3, 3, 300, 2))
    want = O.run_dynamic([{"id": "task", "num_qubits": 3, "num_clbits": 3, "instructions": ins}], 300, seed=2)["task"]
    assert res["counts"] == want


def test_batching_does_not_change_results(backend):
    ins = C.hea_ansatz(5, 2, measure=False) + [{"name": "measure", "qubits": [2], "clbits": [0]},
                                               {"name": "cif", "clbits": [0], "operation": "and", "condition": 1,
                                                "instructions": [{"name": "y", "qubits": [4]}]}] + C.measure_all(5)
    a = backend.execute(dyn_task(ins, 5, 5, 257, 9))
    b = backend.execute(dyn_task(ins, 5, 5, 257, 9, b200_max_batch=1))
    c = backend.execute(dyn_task(ins, 5, 5, 257, 9, b200_max_batch=100))
    d = backend.execute(dyn_task(ins, 5, 5, 257, 9, b200_share_prefix=False))  # no shared-prefix shortcut
    assert a["counts"] == b["counts"] == c["counts"] == d["counts"] and sum(a["counts"].values()) == 257


@pytest.mark.parametrize("maker", [C.hea_teledata_pair, C.hea_telegate_pair])
def test_quantum_communication_joint_register(backend, maker):
    """executor-style call: several tasks, one joint register + comm pair (aer_simulator_adapter.cpp:131-159)"""
    tasks = maker(4)
    shots = 400
    msg = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=33, task_id=t["id"], is_dynamic=True,
                  b200_defer_measurement=False) for t in tasks]
    res = backend.execute(msg)
    assert "ERROR" not in res, res
    assert backend.last_mode() == 2
    want = O.run_dynamic(tasks, shots, seed=33)
    if maker is C.hea_teledata_pair:
        # qrecv ends with a swap, which the engine performs by relabelling qubits: its terminal inverse-CDF then
        # walks the amplitudes in another order than the oracle's -- same distribution, not the same shots
        for t in tasks:
            assert sum(res["id_counts"][t["id"]].values()) == shots
            assert O.tvd(res["id_counts"][t["id"]], want[t["id"]]) < 0.5 * math.sqrt(16 / shots) + 0.06
        # with the shortcut off, every shot is the literal measure sequence and must match the oracle exactly
        for m in msg:
            m["config"]["b200_terminal_sampling"] = False
        res = backend.execute(msg)
        assert res["id_counts"] == O.run_dynamic(tasks, shots, seed=33, terminal_sampling=False)
    else:
        assert res["id_counts"] == want


def test_teledata_teleports(backend):
    theta = 1.1
    a = [{"name": "ry", "qubits": [0], "params": [theta]}, {"name": "qsend", "qubits": [0], "qpus": ["B"]}]
    b = [{"name": "qrecv", "qubits": [0], "qpus": ["A"]}, {"name": "measure", "qubits": [0], "clbits": [0]}]
    shots = 4000
    msg = [C.task(a, 1, shots=shots, seed=1, task_id="A", is_dynamic=True, b200_teledata_corrections="protocol"),
           C.task(b, 1, shots=shots, seed=1, task_id="B", is_dynamic=True)]
    # the documented protocol (X <- comm qubit, Z <- sent qubit) is opt-in and does teleport
    res = backend.execute(msg)
    p1 = res["id_counts"]["B"].get("1", 0) / shots
    assert abs(p1 - math.sin(theta / 2) ** 2) < 4 * math.sqrt(0.25 / shots)
    assert backend.last_mode() == 1  # measurements deferred, corrections as quantum control
    assert res["b200"] == {"mode": "deferred_measurement", "teledata_corrections": "protocol"}
    # DEFAULT = the reference's literal code (aer_simulator_adapter.cpp:644-658: X <- sent qubit, Z <- comm qubit, quirk
    # Q6): same histograms as the reference on the same inputs -- the received qubit reads 1 with probability 1/2
    del msg[0]["config"]["b200_teledata_corrections"]
    res = backend.execute(msg)
    assert res["b200"]["teledata_corrections"] == "reference"
    assert abs(res["id_counts"]["B"].get("1", 0) / shots - 0.5) < 4 * math.sqrt(0.25 / shots)
    msg[0]["config"]["b200_defer_measurement"] = False  # per-shot interpreter: shot-identical to the oracle
    lit = backend.execute(msg)
    want = O.run_dynamic([{"id": "A", "num_qubits": 1, "num_clbits": 1, "instructions": a},
                          {"id": "B", "num_qubits": 1, "num_clbits": 1, "instructions": b}], shots, seed=1)
    assert lit["id_counts"] == want and lit["b200"] == {"mode": "per_shot_batched", "teledata_corrections": "reference"}


def test_classical_send_recv_between_processes(lib_built):
    """SEND/RECV through the channel hooks (ClassicalChannel::send_measure / recv_measure)"""
    be = cb.Backend(0)
    sent, inbox = [], [1, 0, 1, 1, 0, 0, 1, 0]
    be.set_channel(lambda bit, tgt: sent.append((bit, tgt)), lambda org: inbox.pop(0))
    ins = [{"name": "x", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]},
           {"name": "send", "clbits": [0], "qpus": ["peer"]}, {"name": "recv", "clbits": [1], "qpus": ["peer"]},
           {"name": "cif", "clbits": [1], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [1]}]},
           {"name": "measure", "qubits": [1], "clbits": [1]}]
    res = be.execute(C.task(ins, 2, shots=8, seed=0, is_dynamic=True, sending_to=["peer"]))
    assert "ERROR" not in res, res
    assert sent == [(1, "peer")] * 8
    assert res["counts"] == {"11": 4, "01": 4}
    be.close()


# ---- deferred measurement (default for dynamic tasks): one state, feed-forward as quantum control -------------
def _dist(counts):
    tot = sum(counts.values())
    return {k: v / tot for k, v in counts.items()}


def test_deferred_measurement_cif(backend):
    ins = [{"name": "h", "qubits": [0]}, {"name": "ry", "qubits": [2], "params": [0.9]},
           {"name": "measure", "qubits": [0], "clbits": [0]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 1,
            "instructions": [{"name": "x", "qubits": [1]}, {"name": "rx", "qubits": [2], "params": [1.3]},
                             C.unitary_inst(C.haar_unitary(np.random.default_rng(4), 4), [1, 2])]},
           {"name": "cz", "qubits": [0, 2]},                      # a measured qubit may still act as a control
           {"name": "measure", "qubits": [1], "clbits": [1]}, {"name": "measure", "qubits": [2], "clbits": [2]},
           {"name": "measure", "qubits": [0], "clbits": [3]}]     # measured again: same outcome
    shots = 20000
    res = backend.execute(dyn_task(ins, 3, 4, shots, 17, b200_defer_measurement=True))
    assert "ERROR" not in res, res
    assert backend.last_mode() == 1
    assert all(k[0] == k[3] for k in res["counts"])  # clbit 3 repeats clbit 0
    # the oracle restates the deferral rule: same final state, same u_s = U(seed, s) -> the same shots
    assert res["counts"] == O.run_dynamic([{"id": "task", "num_qubits": 3, "num_clbits": 4, "instructions": ins}],
                                          shots, seed=17, defer=True)["task"]
    ref = backend.execute(dyn_task(ins, 3, 4, shots, 18))
    assert backend.last_mode() == 2
    assert O.tvd(res["counts"], ref["counts"]) < 0.03
    want = O.run_dynamic([{"id": "task", "num_qubits": 3, "num_clbits": 4, "instructions": ins}], 4000, seed=5)["task"]
    assert O.tvd(res["counts"], want) < 0.5 * math.sqrt(8 / 4000) + 0.03


@pytest.mark.parametrize("maker", [C.hea_teledata_pair, C.hea_telegate_pair])
def test_deferred_measurement_quantum_communication(backend, maker):
    tasks = maker(4, layers=2)
    shots = 40000
    msg = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=7, task_id=t["id"], is_dynamic=True) for t in tasks]
    res = backend.execute(msg)
    assert "ERROR" not in res, res
    assert backend.last_mode() == 1
    if maker is C.hea_telegate_pair:  # (teledata ends with a relabelled swap: distribution-equal only)
        assert res["id_counts"] == O.run_dynamic(tasks, shots, seed=7, defer=True)
    st = backend.stats()
    for m in msg:
        m["config"]["b200_defer_measurement"] = False
        m["config"]["seed"] = 8
    ref = backend.execute(msg)
    assert backend.last_mode() == 2
    assert backend.stats()["bytes"] > 20 * st["bytes"]  # the per-shot interpreter moves far more data
    for t in tasks:
        assert sum(res["id_counts"][t["id"]].values()) == shots
        assert O.tvd(res["id_counts"][t["id"]], ref["id_counts"][t["id"]]) < 0.03
    want = O.run_dynamic(tasks, 3000, seed=3)
    for t in tasks:
        assert O.tvd(res["id_counts"][t["id"]], want[t["id"]]) < 0.5 * math.sqrt(16 / 3000) + 0.04


@pytest.mark.parametrize("ins", [
    # a measured qubit is driven again
    [{"name": "h", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "h", "qubits": [0]},
     {"name": "measure", "qubits": [0], "clbits": [1]}],
    # reset of a live qubit
    [{"name": "h", "qubits": [0]}, {"name": "cx", "qubits": [0, 1]}, {"name": "reset", "qubits": [0]},
     {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "measure", "qubits": [1], "clbits": [1]}],
    # a two-bit condition on undecided bits
    [{"name": "h", "qubits": [0]}, {"name": "h", "qubits": [1]}, {"name": "measure", "qubits": [0], "clbits": [0]},
     {"name": "measure", "qubits": [1], "clbits": [1]},
     {"name": "cif", "clbits": [0, 1], "operation": "xor", "condition": 1, "instructions": [{"name": "x", "qubits": [1]}]}],
])
def test_deferral_falls_back_when_a_collapse_is_needed(backend, ins):
    res = backend.execute(dyn_task(ins, 2, 2, 500, 21, b200_defer_measurement=True))
    assert "ERROR" not in res, res
    assert backend.last_mode() == 2
    want = O.run_dynamic([{"id": "task", "num_qubits": 2, "num_clbits": 2, "instructions": ins}], 500, seed=21)["task"]
    assert res["counts"] == want


def test_deferred_measurement_reuses_communication_pairs(backend):
    """two teleports through ONE pair with the documented corrections: the first teleport leaves its measured qubits
    disentangled (|+>), so they are reset by a unitary and the second teleport is deferred as well; with two pairs a fresh
    pair is taken.  With the reference's literal corrections (the default, quirk Q6) the measured qubits stay entangled
    with the received one: reusing the pair needs a real collapse, and the engine falls back to the per-shot interpreter."""
    a = [{"name": "ry", "qubits": [0], "params": [0.7]}, {"name": "qsend", "qubits": [0], "qpus": ["B"]},
         {"name": "ry", "qubits": [1], "params": [0.3]}, {"name": "qsend", "qubits": [1], "qpus": ["B"]}]
    b = [{"name": "qrecv", "qubits": [0], "qpus": ["A"]}, {"name": "qrecv", "qubits": [1], "qpus": ["A"]}] + C.measure_all(2)
    shots = 20000
    p_want = {"00": math.cos(0.35) ** 2 * math.cos(0.15) ** 2, "01": math.sin(0.35) ** 2 * math.cos(0.15) ** 2,
              "10": math.cos(0.35) ** 2 * math.sin(0.15) ** 2, "11": math.sin(0.35) ** 2 * math.sin(0.15) ** 2}
    for n_comm, mode in ((2, 1), (4, 1)):
        msg = [C.task(a, 2, shots=shots, seed=3, task_id="A", is_dynamic=True, n_communication_qubits=n_comm,
                      b200_teledata_corrections="protocol"),
               C.task(b, 2, shots=shots, seed=3, task_id="B", is_dynamic=True, n_communication_qubits=n_comm)]
        res = backend.execute(msg)
        assert "ERROR" not in res, res
        assert backend.last_mode() == mode
        got = res["id_counts"]["B"]
        for k, p in p_want.items():
            assert abs(got.get(k, 0) / shots - p) < 5 * math.sqrt(p * (1 - p) / shots) + 1e-3, (n_comm, k)
    # default = the reference's literal corrections: one pair cannot be recycled without a collapse -> per-shot interpreter,
    # shot-identical to the oracle (which restates the literal code)
    msg = [C.task(a, 2, shots=600, seed=3, task_id="A", is_dynamic=True, b200_terminal_sampling=False),
           C.task(b, 2, shots=600, seed=3, task_id="B", is_dynamic=True)]
    res = backend.execute(msg)
    assert backend.last_mode() == 2 and res["b200"]["teledata_corrections"] == "reference"
    tasks = [{"id": "A", "num_qubits": 2, "num_clbits": 2, "instructions": a}, {"id": "B", "num_qubits": 2, "num_clbits": 2, "instructions": b}]
    assert res["id_counts"] == O.run_dynamic(tasks, 600, seed=3, terminal_sampling=False)


def test_deferred_measurement_many_remote_operations(backend):
    """teleports, a re-driven sender qubit and a remote-controlled gate through one pair: all deferred, and the
    shots equal the oracle's deferred shots except for the relabelled swaps (distribution-equal)"""
    from test_oracle import _many_remote_ops
    tasks = _many_remote_ops(2)
    shots = 40000
    msg = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=5, task_id=t["id"], is_dynamic=True,
                  b200_teledata_corrections="protocol") for t in tasks]
    res = backend.execute(msg)
    assert "ERROR" not in res, res
    assert backend.last_mode() == 1
    dist = O.deferred_distribution(tasks, n_comm_qubits=2, teledata_literal=False)
    for tid in ("A", "B"):
        want = {}
        for key, p in dist.items():
            k = dict(key)[tid]
            want[k] = want.get(k, 0.0) + p
        got = res["id_counts"][tid]
        assert sum(got.values()) == shots
        assert 0.5 * sum(abs(got.get(k, 0) / shots - want.get(k, 0.0)) for k in set(got) | set(want)) < 0.02


@pytest.mark.parametrize("which", ["_negated_condition", "_active_reset", "_nested_and_multibit_conditions"])
def test_deferred_measurement_condition_forms(backend, which):
    import test_oracle
    t = getattr(test_oracle, which)()[0]
    shots = 20000
    res = backend.execute(dyn_task(t["instructions"], t["num_qubits"], t["num_clbits"], shots, 4, b200_defer_measurement=True))
    assert "ERROR" not in res, res
    assert backend.last_mode() == 1
    assert res["counts"] == O.run_dynamic([t], shots, seed=4, defer=True)["task"]


def test_fusion_config_keys(backend):
    """Aer's fusion switches are honoured per task: same state, different op count"""
    ins = C.hea_ansatz(6, 3, 7, measure=False) + [{"name": "save_state", "qubits": list(range(6))}] + C.measure_all(6)
    a = backend.execute(C.task(ins, 6, shots=50, seed=1))
    ops_default = backend.stats()["fused_ops"]
    b = backend.execute(C.task(ins, 6, shots=50, seed=1, fusion_enable=False))
    ops_unfused = backend.stats()["fused_ops"]
    c = backend.execute(C.task(ins, 6, shots=50, seed=1, fusion_enable=True, fusion_max_qubit=1))
    for r in (a, b, c):
        assert "ERROR" not in r, r
    sa, sb, sc = (np.array(r["statevector"]) for r in (a, b, c))
    assert np.abs(sa - sb).max() < 1e-12 and np.abs(sa - sc).max() < 1e-12
    assert ops_unfused > ops_default
    assert backend.execute(C.task(ins, 6, shots=50, seed=1))["counts"] == a["counts"]  # the switch does not stick


def test_qasm2_text_runs_on_the_engine(backend):
    """OpenQASM 2 -> circuit JSON (cb200_qasm2_to_json) -> wire task -> counts"""
    from test_qasm2 import GHZ
    circ = cb.qasm2_to_json(GHZ)
    res = backend.execute(C.task(circ["instructions"], circ["num_qubits"], num_clbits=circ["num_clbits"], shots=500, seed=3))
    assert "ERROR" not in res, res
    assert set(res["counts"]) == {"0000", "1111"} and sum(res["counts"].values()) == 500
    assert res["counts"] == O.run_static(circ["instructions"], 4, 4, shots=500, seed=3)[1]


def test_out_of_band_amplitudes_and_marginals_through_the_wire_format(backend):
    """config keys b200_amplitude_indices / b200_marginal_qubits, and save_state above 20 qubits
    (consumers: cunqa/result.py:151-187, src/utils/probabilities/process_counts.hpp:71-131)"""
    n = 22
    ins = C.quantum_volume(n, depth=3, seed=12)
    ref, _ = O.run_static(ins, n, backend="c")
    idx = [0, 5, (1 << n) - 1, 123456]
    qs = [3, 21, 0, 11]
    t = C.task(ins + [{"name": "save_state", "qubits": list(range(n))}] + C.measure_all(n), n, shots=300, seed=2,
               b200_amplitude_indices=idx, b200_marginal_qubits=qs)
    res = backend.execute(t)
    assert "ERROR" not in res, res
    got = np.array([complex(*z) for z in res["b200_amplitudes"]])
    assert np.abs(got - ref.amplitudes()[idx]).max() < 1e-12
    p = np.abs(ref.amplitudes()) ** 2
    i = np.arange(1 << n)
    j = sum(((i >> q) & 1) << m for m, q in enumerate(qs))
    assert np.abs(np.array(res["b200_marginal"]) - np.bincount(j, weights=p, minlength=16)).max() < 1e-12
    # 22 qubits: the statevector comes back sampled -- the amplitudes of the basis states the shots landed on
    sam = res["statevector_sampled"]
    assert "statevector" not in res and 1 <= len(sam["indices"]) <= 300
    a = np.array([complex(*z) for z in sam["amplitudes"]])
    assert np.abs(a - ref.amplitudes()[sam["indices"]]).max() < 1e-12
    assert {int(k, 2) for k in res["counts"]} == set(sam["indices"])
    # up to 20 qubits the reference's shape: the whole vector under "statevector"
    small = backend.execute(C.task(C.ghz(3, measure=False) + [{"name": "save_state", "qubits": [0, 1, 2]}] + C.measure_all(3), 3, shots=10, seed=1))
    v = np.array([complex(*z) for z in small["statevector"]])
    assert abs(v[0] - 2 ** -0.5) < 1e-15 and abs(v[7] - 2 ** -0.5) < 1e-15 and np.abs(v[1:7]).max() < 1e-15
