"""GPU (-m gpu, >= 2 GPUs): ONE state sharded over the GPUs of a box by ONE process -- the mode the plugin uses
(config.device.target_devices = [0, 1, ...], /root/reference/src/utils/helpers/net_functions.hpp:198-222; one process
per vQPU, src/cli/setup_qpus.cpp:69-82).  Everything goes through the C ABI: cb200_state_create_sharded for the
state-level checks, cb200_backend_create_multi + cb200_backend_execute for whole tasks."""
import math

import numpy as np
import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O
from conftest import random_circuit

pytestmark = pytest.mark.gpu


def _devices():
    n = cb.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    return list(range(4 if n >= 4 else 2))


def test_qft_and_random_circuit_amplitudes(lib_built):
    devs = _devices()
    g = len(devs).bit_length() - 1
    n = 14 + g
    x = C.qft_x(n, seed=3)
    with cb.StateVector(n, devices=devs) as sv:
        sv.run(C.qft(n, x))
        idx = np.arange(1 << n, dtype=np.uint64)
        err = np.abs(sv.amplitudes(idx) - O.qft_amplitudes(n, x, idx)).max()
        assert err < 1e-10 and err * 2 ** (n / 2) < 1e-9
        assert abs(sv.norm()[0] - 1) < 1e-10
        assert sv.dist_stats()["world"] == len(devs) and sv.dist_stats()["swaps"] >= g
    n = 13 + g
    ins = random_circuit(n, 150, 21, max_unitary=4)
    ref, _ = O.run_static(ins, n, backend="c")
    with cb.StateVector(n, devices=devs) as sv:
        sv.run(ins)
        assert np.abs(sv.amplitudes() - ref.amplitudes()).max() < 1e-10
        for q in (0, n - 1, n // 2):
            assert abs(sv.prob1(q)[0] - ref.prob1(q)) < 1e-10
        # marginals over local and global qubits (global ones are constant per shard)
        qs = [n - 1, 2, n - 2, 7]
        p = np.abs(ref.amplitudes()) ** 2
        i = np.arange(1 << n)
        j = sum(((i >> q) & 1) << m for m, q in enumerate(qs))
        assert np.abs(sv.probabilities(qs) - np.bincount(j, weights=p, minlength=16)).max() < 1e-12
        shots = sv.sample(4000, 9)
        emp = np.bincount(shots.astype(np.int64), minlength=1 << n) / 4000.0
        assert 0.5 * np.abs(emp - p).sum() < 0.5 * math.sqrt((1 << n) / 4000.0) + 0.05


def test_global_controls_drop_ops_on_some_ranks(lib_built):
    """ops controlled by a global qubit exist on some shards only; the exchanges must still line up"""
    devs = _devices()
    g = len(devs).bit_length() - 1
    n = 13 + g
    top = n - 1
    ins = [{"name": "h", "qubits": [q]} for q in range(n)]
    ins += [{"name": "cx", "qubits": [top, top - 1 if g > 1 else 3]}] + [{"name": "cx", "qubits": [top, q]} for q in range(2, 10)]
    ins += [{"name": "h", "qubits": [top - 1 if g > 1 else top]}, {"name": "ccx", "qubits": [top, 0, 5]}, {"name": "ry", "qubits": [top], "params": [0.4]}]
    ref, _ = O.run_static(ins, n, backend="c")
    with cb.StateVector(n, devices=devs) as sv:
        sv.run(ins)
        assert np.abs(sv.amplitudes() - ref.amplitudes()).max() < 1e-10


def test_measure_reset_and_set_amplitudes(lib_built):
    devs = _devices()
    g = len(devs).bit_length() - 1
    n = 12 + g
    ins = C.quantum_volume(n, depth=3, seed=5)
    ref, _ = O.run_static(ins, n, backend="c")
    with cb.StateVector(n, devices=devs) as sv:
        sv.run(ins)
        for step, (q, rst) in enumerate([(1, False), (n - 1, False), (n - 2, True), (0, True)]):
            want = ref.measure(q, 5, step) if not rst else None
            if rst:
                bit = ref.measure(q, 5, step)
                if bit:
                    O.apply_gate(ref, {"name": "x", "qubits": [q]})
                sv.measure(q, seed=5, counters=[step], reset=True)
            else:
                assert sv.measure(q, seed=5, counters=[step])[0] == want
            assert np.abs(sv.amplitudes() - ref.amplitudes()).max() < 1e-10
        rng = np.random.default_rng(2)
        v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        v /= np.linalg.norm(v)
        sv.set_amplitudes(v)
        assert np.abs(sv.amplitudes() - v).max() < 1e-15


def test_backend_execute_on_target_devices(lib_built):
    """static task, {"params"} update and a dynamic task through cb200_backend_execute on a sharded register"""
    devs = _devices()
    g = len(devs).bit_length() - 1
    n = 13 + g
    be = cb.Backend(devices=devs)
    shots = 3000
    ins = C.hea_ansatz(n, layers=2, seed=4)
    task = C.task(ins, n, shots=shots, seed=8, b200_shard_min_qubits=n)
    task["config"]["device"]["target_devices"] = devs
    res = be.execute(task)
    assert "ERROR" not in res, res
    ref, want = O.run_static(ins, n, shots=shots, seed=8, backend="c")
    # (the engine's inverse CDF walks the amplitudes in the physical, rank-major order: same distribution, other shots)
    assert sum(res["counts"].values()) == shots and want
    p = np.abs(ref.amplitudes()) ** 2
    top = sorted(res["counts"], key=res["counts"].get)[-1]
    assert abs(res["counts"][top] / shots - p[int(top, 2)]) < 5 * math.sqrt(p[int(top, 2)] / shots) + 1e-3
    idx = np.arange(1 << n)
    for q in range(n):   # every single-qubit marginal within 5 sigma of the exact one (clbit q <-> qubit q)
        p1 = p[((idx >> q) & 1) == 1].sum()
        emp = sum(c for k, c in res["counts"].items() if k[n - 1 - q] == "1") / shots
        assert abs(emp - p1) < 5 * math.sqrt(max(p1 * (1 - p1), 1e-4) / shots), q
    # parameter update re-runs the cached task (quantum_task.cpp:39-108)
    newp = list(np.random.default_rng(1).uniform(0, 6.28, 4 * n))
    res = be.execute({"params": newp, "shots": 500})
    assert "ERROR" not in res and sum(res["counts"].values()) == 500
    # GHZ over every qubit incl. the global ones, measured mid-circuit with feed-forward: deferred mode and per-shot mode
    dyn = [{"name": "h", "qubits": [n - 1]}] + [{"name": "cx", "qubits": [n - 1, q]} for q in range(n - 1)]
    dyn += [{"name": "measure", "qubits": [n - 1], "clbits": [0]},
            {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [0]}]},
            {"name": "measure", "qubits": [0], "clbits": [1]}, {"name": "measure", "qubits": [n - 2], "clbits": [2]}]
    for defer in (True, False):
        t = C.task(dyn, n, 3, shots=200 if defer else 24, seed=3, is_dynamic=True, b200_shard_min_qubits=n, b200_defer_measurement=defer)
        res = be.execute(t)
        assert "ERROR" not in res, res
        assert be.last_mode() == (1 if defer else 2)
        assert set(res["counts"]) <= {"000", "101"}          # c0 = c2 = GHZ outcome, c1 = q0 flipped back to 0 when it was 1
    be.close()


def test_quantum_communication_group_on_a_sharded_register(lib_built):
    """two vQPU tasks + communication pair as ONE joint register sharded over the GPUs (the QC executor's message through
    cb200_backend_execute with a device list): telegate, deferred and per-shot"""
    devs = _devices()
    tasks = C.hea_telegate_pair(6, layers=1, seed=2)
    n_joint = 14
    be = cb.Backend(devices=devs)
    for defer, shots in ((True, 2000), (False, 16)):
        # (per-shot mode: the literal measure-by-measure sequence, so that engine and oracle draw the same stream -- the
        # terminal inverse-CDF shortcut walks a sharded state in another order than the oracle's)
        msg = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=4, task_id=t["id"], is_dynamic=True,
                      b200_shard_min_qubits=n_joint, b200_defer_measurement=defer, b200_terminal_sampling=False) for t in tasks]
        res = be.execute(msg)
        assert "ERROR" not in res, res
        assert be.last_mode() == (1 if defer else 2)
        want = O.run_dynamic(tasks, shots, seed=4, defer=defer, terminal_sampling=False)
        for tid in want:
            assert sum(res["id_counts"][tid].values()) == shots
            if not defer:
                assert res["id_counts"][tid] == want[tid]          # same measurement stream, shot for shot
            else:
                assert O.tvd(res["id_counts"][tid], want[tid]) < 0.5 * math.sqrt(2 ** 6 / shots) + 0.05
    # out-of-band amplitudes and marginals of a sharded register through the wire format
    n = 13 + len(devs).bit_length() - 1
    ins = C.quantum_volume(n, depth=2, seed=6)
    ref, _ = O.run_static(ins, n, backend="c")
    idx, qs = [0, 3, (1 << n) - 1], [n - 1, 0, 5]
    r = be.execute(C.task(ins + C.measure_all(n), n, shots=50, seed=1, b200_shard_min_qubits=n, b200_amplitude_indices=idx, b200_marginal_qubits=qs))
    assert "ERROR" not in r, r
    assert np.abs(np.array([complex(*z) for z in r["b200_amplitudes"]]) - ref.amplitudes()[idx]).max() < 1e-12
    p = np.abs(ref.amplitudes()) ** 2
    i = np.arange(1 << n)
    j = sum(((i >> q) & 1) << m for m, q in enumerate(qs))
    assert np.abs(np.array(r["b200_marginal"]) - np.bincount(j, weights=p, minlength=8)).max() < 1e-12
    be.close()
