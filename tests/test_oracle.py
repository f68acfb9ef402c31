"""CPU: pin the oracle (C restatement + numpy twin) against analytic known answers and each other.

The reference holds no numeric fixtures for this path (parity unpinned), so the
anchors are closed forms: GHZ, QFT|x>, U^dagger U = 1, teleportation fidelity.
"""
import math

import numpy as np
import pytest

from cunqa_b200 import circuits as C
from oracle import oracle_np as O
from conftest import random_circuit


@pytest.mark.parametrize("backend", ["numpy", "c"])
@pytest.mark.parametrize("n", [1, 2, 5, 11])
def test_qft_closed_form(lib_built, backend, n):
    x = C.qft_x(n, seed=n)
    st, _ = O.run_static(C.qft(n, x), n, backend=backend)
    ref = O.qft_amplitudes(n, x, np.arange(1 << n))
    assert np.abs(st.amplitudes() - ref).max() < 1e-13


@pytest.mark.parametrize("backend", ["numpy", "c"])
def test_ghz(lib_built, backend):
    n = 9
    st, counts = O.run_static(C.ghz(n), n, shots=2000, seed=11, backend=backend)
    a = st.amplitudes()
    assert abs(a[0] - 1 / math.sqrt(2)) < 1e-14 and abs(a[-1] - 1 / math.sqrt(2)) < 1e-14
    assert np.abs(a[1:-1]).max() < 1e-15
    assert set(counts) <= {"0" * n, "1" * n}
    assert abs(counts["0" * n] / 2000 - 0.5) < 4 * math.sqrt(0.25 / 2000)


@pytest.mark.parametrize("seed", range(6))
def test_numpy_twin_equals_c(lib_built, seed):
    n = 3 + seed
    ins = random_circuit(n, 120, seed)
    a, _ = O.run_static(ins, n, backend="numpy")
    b, _ = O.run_static(ins, n, backend="c")
    assert np.abs(a.amplitudes() - b.amplitudes()).max() < 1e-13
    assert abs(a.norm() - 1) < 1e-12


def test_inverse_roundtrip(lib_built):
    n = 10
    ins = C.quantum_volume(n, depth=6, seed=5)
    st, _ = O.run_static(ins + C.inverse(ins), n, backend="c")
    a = st.amplitudes()
    assert abs(a[0] - 1) < 1e-12 and np.abs(a[1:]).max() < 1e-12


def test_sampling_same_stream_both_backends(lib_built):
    n = 8
    ins = C.quantum_volume(n, depth=4, seed=1, measure=True)
    a, ca = O.run_static(ins, n, shots=500, seed=42, backend="numpy")
    b, cb = O.run_static(ins, n, shots=500, seed=42, backend="c")
    assert ca == cb  # same splitmix stream, same inverse-CDF rule
    p = np.abs(a.amplitudes()) ** 2
    emp = np.zeros(1 << n)
    for k, v in ca.items():
        emp[int(k, 2)] = v / 500
    assert 0.5 * np.abs(emp - p).sum() < 0.5 * math.sqrt((1 << n) / 500) + 0.05


def test_measure_collapse(lib_built):
    for backend in ("numpy", "c"):
        st = O.make_state(3, backend)
        st.apply_matrix([0], O._FIXED_1Q["h"])
        st.apply_matrix([1], O._FIXED_1Q["x"], [0])
        out = st.measure(0, seed=3, counter=0)
        a = st.amplitudes()
        assert abs(abs(a[3 if out else 0]) - 1) < 1e-14 and abs(st.norm() - 1) < 1e-14


def test_bit_order_clbit0_rightmost():
    # aer_helpers.hpp:175-179: clbit 0 is the rightmost character
    assert O.bits_to_string({0: 1}, 3) == "001"
    assert O.bits_to_string({2: 1}, 3) == "100"


def test_teledata_moves_the_state(lib_built):
    """QSEND/QRECV restatement with the DOCUMENTED corrections (X <- comm qubit, Z <- sent qubit): B's qubit 0 must end
    up in the state A prepared.  The reference's literal code has the two swapped (quirk Q6, the oracle's default):
    the received qubit is then maximally mixed in the Z basis whatever A prepared."""
    theta = 1.1
    a = [{"name": "ry", "qubits": [0], "params": [theta]}, {"name": "qsend", "qubits": [0], "qpus": ["B"]}]
    b = [{"name": "qrecv", "qubits": [0], "qpus": ["A"]}, {"name": "measure", "qubits": [0], "clbits": [0]}]
    tasks = [{"id": "A", "num_qubits": 1, "num_clbits": 1, "instructions": a},
             {"id": "B", "num_qubits": 1, "num_clbits": 1, "instructions": b}]
    shots = 4000
    res = O.run_dynamic(tasks, shots, seed=9, teledata_literal=False)
    p1 = res["B"].get("1", 0) / shots
    assert abs(p1 - math.sin(theta / 2) ** 2) < 4 * math.sqrt(0.25 / shots)
    # the reference's literal sequence (aer_simulator_adapter.cpp:644-658): exact distribution by branch enumeration
    lit = O.literal_distribution(tasks, teledata_literal=True)
    assert abs(sum(p for k, p in lit.items() if dict(k)["B"] == "1") - 0.5) < 1e-12
    doc = O.literal_distribution(tasks, teledata_literal=False)
    assert abs(sum(p for k, p in doc.items() if dict(k)["B"] == "1") - math.sin(theta / 2) ** 2) < 1e-12


def test_telegate_is_a_remote_cx(lib_built):
    """EXPOSE/RCONTROL restatement: remote cx from A's |+> onto B's |0> gives a Bell pair."""
    a = [{"name": "h", "qubits": [0]}, {"name": "expose", "qubits": [0], "qpus": ["B"]},
         {"name": "measure", "qubits": [0], "clbits": [0]}]
    b = [{"name": "rcontrol", "qpus": ["A"], "instructions": [{"name": "cx", "qubits": [-1, 0]}]},
         {"name": "measure", "qubits": [0], "clbits": [0]}]
    tasks = [{"id": "A", "num_qubits": 1, "num_clbits": 1, "instructions": a},
             {"id": "B", "num_qubits": 1, "num_clbits": 1, "instructions": b}]
    # joint outcomes must be perfectly correlated: run shot by shot with the same seed and compare
    for shot_seed in range(40):
        r = O.run_dynamic(tasks, 1, seed=shot_seed)
        assert list(r["A"]) == list(r["B"])


def test_cif_feed_forward(lib_built):
    ins = [{"name": "h", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 1,
            "instructions": [{"name": "x", "qubits": [1]}]},
           {"name": "measure", "qubits": [1], "clbits": [1]}]
    res = O.run_dynamic([{"id": "t", "num_qubits": 2, "num_clbits": 2, "instructions": ins}], 200, seed=4)
    assert set(res["t"]) <= {"00", "11"} and len(res["t"]) == 2


def test_terminal_sampling_is_distribution_equivalent(lib_built):
    """the engine's shortcut (one |a|^2 draw when only measurements remain) vs the reference's literal
    sequence of apply_measure calls: same distribution"""
    ins = [{"name": "h", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]},
           {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "ry", "qubits": [1], "params": [1.0]}]},
           {"name": "rx", "qubits": [2], "params": [0.8]}, {"name": "cx", "qubits": [2, 1]},
           {"name": "measure", "qubits": [1], "clbits": [1]}, {"name": "measure", "qubits": [2], "clbits": [2]}]
    task = [{"id": "t", "num_qubits": 3, "num_clbits": 3, "instructions": ins}]
    shots = 3000
    a = O.run_dynamic(task, shots, seed=1, terminal_sampling=True)["t"]
    b = O.run_dynamic(task, shots, seed=2, terminal_sampling=False)["t"]
    assert O.tvd(a, b) < 0.5 * math.sqrt(8 / shots) + 0.03


def test_block_fusion_for_the_cpu_baseline(lib_built):
    """oracle-side greedy <=4-qubit fusion (bench.py's CPU baseline) is an exact regrouping"""
    n = 9
    ins = C.quantum_volume(n, depth=5, seed=3) + random_circuit(n, 40, 2)
    ref, _ = O.run_static(ins, n, backend="numpy")
    blocks = O.fuse_blocks(ins, kmax=4)
    assert len(blocks) < len(ins) and all(len(q) <= 5 for q, _ in blocks)
    st = O.COracleState(n)
    for qs, m in blocks:
        st.apply_matrix(qs, m)
    assert np.abs(st.amplitudes() - ref.amplitudes()).max() < 1e-12


# ---- deferred measurement: the rule the engine applies by default to dynamic tasks, proven on small circuits -----
def _cif_circuit():
    return [{"id": "task", "num_qubits": 3, "num_clbits": 4, "instructions": [
        {"name": "h", "qubits": [0]}, {"name": "ry", "qubits": [2], "params": [0.9]},
        {"name": "measure", "qubits": [0], "clbits": [0]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 1,
         "instructions": [{"name": "x", "qubits": [1]}, {"name": "rx", "qubits": [2], "params": [1.3]},
                          C.unitary_inst(C.haar_unitary(np.random.default_rng(4), 4), [1, 2]),
                          {"name": "gp", "qubits": [], "params": [0.4]}]},
        {"name": "cz", "qubits": [0, 2]},
        {"name": "measure", "qubits": [1], "clbits": [1]}, {"name": "measure", "qubits": [2], "clbits": [2]},
        {"name": "measure", "qubits": [0], "clbits": [3]}]}]


def _negated_condition():
    return [{"id": "task", "num_qubits": 3, "num_clbits": 3, "instructions": [
        {"name": "ry", "qubits": [0], "params": [1.0]}, {"name": "h", "qubits": [2]}, {"name": "measure", "qubits": [0], "clbits": [0]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 0,
         "instructions": [{"name": "x", "qubits": [1]}, {"name": "crz", "qubits": [2, 1], "params": [0.6]}, {"name": "h", "qubits": [2]}]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "ry", "qubits": [1], "params": [0.5]}]},
        {"name": "measure", "qubits": [1], "clbits": [1]}, {"name": "measure", "qubits": [2], "clbits": [2]}]}]


def _active_reset():
    """measure + conditional X (the usual active reset), the qubit re-used afterwards"""
    return [{"id": "task", "num_qubits": 2, "num_clbits": 3, "instructions": [
        {"name": "ry", "qubits": [0], "params": [1.2]}, {"name": "cx", "qubits": [0, 1]},
        {"name": "measure", "qubits": [0], "clbits": [0]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [0]}]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "ry", "qubits": [1], "params": [0.5]}]},
        {"name": "measure", "qubits": [0], "clbits": [2]}, {"name": "measure", "qubits": [1], "clbits": [1]}]}]


def _nested_and_multibit_conditions():
    return [{"id": "task", "num_qubits": 4, "num_clbits": 4, "instructions": [
        {"name": "ry", "qubits": [0], "params": [1.0]}, {"name": "ry", "qubits": [1], "params": [2.0]}, {"name": "h", "qubits": [2]},
        {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "measure", "qubits": [1], "clbits": [1]},
        {"name": "cif", "clbits": [0, 1], "operation": "and", "condition": 1,
         "instructions": [{"name": "ry", "qubits": [3], "params": [0.9]}, {"name": "gp", "qubits": [], "params": [0.3]}]},
        {"name": "cif", "clbits": [0], "operation": "and", "condition": 0, "instructions": [
            {"name": "cif", "clbits": [1], "operation": "and", "condition": 1,
             "instructions": [{"name": "cx", "qubits": [2, 3]}, {"name": "diagonal", "qubits": [2], "matrix": [[[1, 0], [0, 1]]]}]},
            {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [3]}]}]},  # contradiction: never
        {"name": "h", "qubits": [2]},
        {"name": "measure", "qubits": [2], "clbits": [2]}, {"name": "measure", "qubits": [3], "clbits": [3]}]}]


def _two_way_traffic():
    """teledata A->B, then a classically communicated outcome B->A conditions a gate on A"""
    a = C.hea_ansatz(2, 1, 5, measure=False) + [{"name": "qsend", "qubits": [1], "qpus": ["B"]},
                                                 {"name": "recv", "clbits": [1], "qpus": ["B"]},
                                                 {"name": "cif", "clbits": [1], "operation": "and", "condition": 1,
                                                  "instructions": [{"name": "ry", "qubits": [0], "params": [0.8]}]},
                                                 {"name": "measure", "qubits": [0], "clbits": [0]}]
    b = C.hea_ansatz(2, 1, 6, measure=False) + [{"name": "qrecv", "qubits": [0], "qpus": ["A"]},
                                                 {"name": "cx", "qubits": [0, 1]},
                                                 {"name": "measure", "qubits": [1], "clbits": [1]},
                                                 {"name": "send", "clbits": [1], "qpus": ["A"]},
                                                 {"name": "measure", "qubits": [0], "clbits": [0]}]
    return [{"id": "A", "num_qubits": 2, "num_clbits": 2, "instructions": a},
            {"id": "B", "num_qubits": 2, "num_clbits": 2, "instructions": b}]


@pytest.mark.parametrize("maker", [_cif_circuit, lambda: C.hea_teledata_pair(2, 1), lambda: C.hea_telegate_pair(2, 1),
                                   lambda: C.hea_telegate_pair(3, 2), _two_way_traffic, _negated_condition, _active_reset,
                                   _nested_and_multibit_conditions])
def test_deferred_measurement_has_the_literal_distribution(maker):
    """exact: joint distribution of all classical outcomes, deferred run vs every branch of the per-shot semantics"""
    tasks = maker()
    lit = O.literal_distribution(tasks)
    dfr = O.deferred_distribution(tasks)
    assert abs(sum(lit.values()) - 1) < 1e-12 and abs(sum(dfr.values()) - 1) < 1e-12
    for key in set(lit) | set(dfr):
        assert abs(lit.get(key, 0.0) - dfr.get(key, 0.0)) < 1e-12, key
    assert len(lit) > 2


@pytest.mark.parametrize("ins", [
    [{"name": "h", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "h", "qubits": [0]},
     {"name": "measure", "qubits": [0], "clbits": [1]}],
    [{"name": "h", "qubits": [0]}, {"name": "cx", "qubits": [0, 1]}, {"name": "reset", "qubits": [0]},
     {"name": "measure", "qubits": [1], "clbits": [1]}],
    [{"name": "h", "qubits": [0]}, {"name": "h", "qubits": [1]}, {"name": "measure", "qubits": [0], "clbits": [0]},
     {"name": "measure", "qubits": [1], "clbits": [1]},
     {"name": "cif", "clbits": [0, 1], "operation": "xor", "condition": 1, "instructions": [{"name": "x", "qubits": [1]}]}],
])
def test_deferral_is_refused_when_a_collapse_is_needed(ins):
    tasks = [{"id": "task", "num_qubits": 2, "num_clbits": 2, "instructions": ins}]
    with pytest.raises(O.DeferAbort):
        O.run_dynamic(tasks, 10, seed=1, defer=True)
    counts, mode = O.run_dynamic_auto(tasks, 50, seed=1)
    assert mode == "per-shot" and counts == O.run_dynamic(tasks, 50, seed=1)


def _many_remote_ops(n_comm):
    """two teleports A->B and a remote-controlled gate B->A through the same communication pair(s)"""
    a = [{"name": "ry", "qubits": [0], "params": [0.7]}, {"name": "qsend", "qubits": [0], "qpus": ["B"]},
         {"name": "ry", "qubits": [1], "params": [0.3]}, {"name": "h", "qubits": [0]},   # the sent qubit is driven again
         {"name": "qsend", "qubits": [1], "qpus": ["B"]},
         {"name": "rcontrol", "qpus": ["B"], "instructions": [{"name": "cx", "qubits": [-1, 0]}]}] + C.measure_all(2)
    b = [{"name": "qrecv", "qubits": [0], "qpus": ["A"]}, {"name": "qrecv", "qubits": [1], "qpus": ["A"]},
         {"name": "cx", "qubits": [0, 1]}, {"name": "expose", "qubits": [1], "qpus": ["A"]}] + C.measure_all(2)
    return [{"id": "A", "num_qubits": 2, "num_clbits": 2, "instructions": a}, {"id": "B", "num_qubits": 2, "num_clbits": 2, "instructions": b}]


@pytest.mark.parametrize("n_comm", [2, 4])
def test_deferred_measurement_recycles_disentangled_qubits(n_comm):
    """a completed teleportation / remote gate leaves its measured qubits in |+>, disentangled: they are reset by a
    unitary, so a communication pair (and the sender's qubit) can be used again without a collapse.  This holds for the
    documented teledata corrections; the reference's literal ones (quirk Q6) leave the measured qubits entangled with
    the received qubit, so a pair that is used AGAIN needs a real collapse: the deferred run gives up (per-shot mode)."""
    tasks = _many_remote_ops(n_comm)
    if n_comm == 2:
        with pytest.raises(O.DeferAbort):
            O.deferred_distribution(tasks, n_comm_qubits=n_comm, teledata_literal=True)
    dfr = O.deferred_distribution(tasks, n_comm_qubits=n_comm, teledata_literal=False)   # must not raise DeferAbort
    lit = O.literal_distribution(tasks, n_comm_qubits=n_comm, teledata_literal=False)
    assert abs(sum(lit.values()) - 1) < 1e-12 and abs(sum(dfr.values()) - 1) < 1e-12 and len(lit) > 4
    for key in set(lit) | set(dfr):
        assert abs(lit.get(key, 0.0) - dfr.get(key, 0.0)) < 1e-12, key


def test_deferred_reset_of_a_disentangled_qubit():
    ins = [{"name": "ry", "qubits": [0], "params": [1.1]}, {"name": "h", "qubits": [1]}, {"name": "cx", "qubits": [1, 2]},
           {"name": "reset", "qubits": [0]}, {"name": "ry", "qubits": [0], "params": [0.4]}, {"name": "cx", "qubits": [0, 1]}] + C.measure_all(3)
    tasks = [{"id": "task", "num_qubits": 3, "num_clbits": 3, "instructions": ins}]
    dfr, lit = O.deferred_distribution(tasks), O.literal_distribution(tasks)
    for key in set(lit) | set(dfr):
        assert abs(lit.get(key, 0.0) - dfr.get(key, 0.0)) < 1e-12, key
