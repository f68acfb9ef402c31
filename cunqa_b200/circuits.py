"""Synthetic workload generators in CUNQA wire format (the `instructions` list of a quantum task).

Instruction shapes follow /root/reference/cunqa/circuit/core.py (e.g. `unitary` :2138-2144,
`cp` :1540-1553) and the canonical message in /root/reference/examples/cpp/simplest_example.cpp:8-28.
The circuits are the ones BASELINE.md section 4 names: GHZ, QFT, random quantum-volume, and a
hardware-efficient VQE ansatz (optionally split over two vQPUs with teledata / telegate).
"""
from __future__ import annotations

import math

import numpy as np


def task(instructions, num_qubits, num_clbits=None, shots=1024, seed=None, task_id="task",
         is_dynamic=False, sending_to=(), device=0, **config):
    """Build the wire message `QuantumTask::update_circuit` parses (src/quantum_task.cpp:18-36)."""
    cfg = {
        "shots": shots, "method": "statevector", "avoid_parallelization": False,
        "num_clbits": num_qubits if num_clbits is None else num_clbits, "num_qubits": num_qubits,
        "device": {"device_name": "GPU", "target_devices": [device]},
    }
    if seed is not None:
        cfg["seed"] = seed
    cfg.update(config)
    return {"id": task_id, "config": cfg, "instructions": list(instructions),
            "sending_to": list(sending_to), "is_dynamic": bool(is_dynamic)}


def measure_all(n):
    return [{"name": "measure", "qubits": [q], "clbits": [q]} for q in range(n)]


def ghz(n, measure=True):
    ins = [{"name": "h", "qubits": [0]}]
    ins += [{"name": "cx", "qubits": [i, i + 1]} for i in range(n - 1)]
    return ins + (measure_all(n) if measure else [])


def qft(n, x=0, swaps=True, measure=False):
    """X-prep of |x>, then for j=n-1..0: h j; cp(pi/2^(j-k)) k,j for k<j; then floor(n/2) swaps."""
    ins = [{"name": "x", "qubits": [q]} for q in range(n) if (x >> q) & 1]
    for j in range(n - 1, -1, -1):
        ins.append({"name": "h", "qubits": [j]})
        for k in range(j - 1, -1, -1):
            ins.append({"name": "cp", "qubits": [k, j], "params": [math.pi / (1 << (j - k))]})
    if swaps:
        ins += [{"name": "swap", "qubits": [i, n - 1 - i]} for i in range(n // 2)]
    return ins + (measure_all(n) if measure else [])


def qft_x(n, seed=7):
    return int(np.random.default_rng(seed).integers(0, 1 << n, dtype=np.uint64)) if n < 64 else 0


def haar_unitary(rng, dim):
    z = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) / math.sqrt(2)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def _wire_matrix(m):
    return [[[float(v.real), float(v.imag)] for v in row] for row in np.asarray(m)]


def unitary_inst(m, qubits):
    return {"name": "unitary", "qubits": list(map(int, qubits)), "matrix": [_wire_matrix(m)]}


def quantum_volume(n, depth=None, seed=2024, measure=False):
    """depth layers; each: seeded permutation of the qubits, floor(n/2) Haar SU(4) `unitary` gates."""
    depth = n if depth is None else depth
    rng = np.random.default_rng(seed)
    ins = []
    for _ in range(depth):
        perm = rng.permutation(n)
        for i in range(n // 2):
            ins.append(unitary_inst(haar_unitary(rng, 4), [perm[2 * i], perm[2 * i + 1]]))
    return ins + (measure_all(n) if measure else [])


def inverse(instructions):
    """inverse circuit for gates that are `unitary`, or self-/param-invertible named gates."""
    out = []
    for inst in reversed(instructions):
        name = inst["name"]
        if name == "unitary":
            a = np.asarray(inst["matrix"][0], dtype=float)
            m = a[..., 0] + 1j * a[..., 1]
            out.append(unitary_inst(m.conj().T, inst["qubits"]))
        elif name in ("h", "x", "y", "z", "cx", "cz", "swap", "ccx", "cswap"):
            out.append(dict(inst))
        elif name in ("rx", "ry", "rz", "p", "cp", "crx", "cry", "crz", "rzz", "rxx", "ryy", "u1", "cu1"):
            out.append({**inst, "params": [-inst["params"][0]]})
        elif name in ("s", "sdg", "t", "tdg", "sx", "sxdg"):
            flip = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t", "sx": "sxdg", "sxdg": "sx"}
            out.append({**inst, "name": flip[name]})
        else:
            raise ValueError(f"no inverse rule for {name}")
    return out


def hea_ansatz(n, layers=2, seed=99, params=None, measure=True):
    """hardware-efficient ansatz: per layer ry+rz on every qubit and a cx ladder."""
    rng = np.random.default_rng(seed)
    if params is None:
        params = rng.uniform(0, 2 * math.pi, size=2 * n * layers)
    params = list(map(float, params))
    ins, p = [], 0
    for _ in range(layers):
        for q in range(n):
            ins.append({"name": "ry", "qubits": [q], "params": [params[p]]})
            ins.append({"name": "rz", "qubits": [q], "params": [params[p + 1]]})
            p += 2
        ins += [{"name": "cx", "qubits": [q, q + 1]} for q in range(n - 1)]
    return ins + (measure_all(n) if measure else [])


def hea_teledata_pair(n_each, layers=1, seed=99, ids=("qpuA", "qpuB")):
    """Two vQPU tasks of n_each qubits running an HEA layer each, then teleporting A's last
    qubit onto B's qubit 0 (qsend/qrecv), a feed-forward-corrected cx on B, and measure-all."""
    a = hea_ansatz(n_each, layers, seed, measure=False)
    b = hea_ansatz(n_each, layers, seed + 1, measure=False)
    a.append({"name": "qsend", "qubits": [n_each - 1], "qpus": [ids[1]]})
    b.append({"name": "qrecv", "qubits": [0], "qpus": [ids[0]]})
    b.append({"name": "cx", "qubits": [0, 1]})
    a += measure_all(n_each)
    b += measure_all(n_each)
    return [
        {"id": ids[0], "num_qubits": n_each, "num_clbits": n_each, "instructions": a},
        {"id": ids[1], "num_qubits": n_each, "num_clbits": n_each, "instructions": b},
    ]


def hea_telegate_pair(n_each, layers=1, seed=99, ids=("qpuA", "qpuB")):
    """Same as above but with a remote-controlled gate: A exposes its last qubit, B applies
    cx(-1 -> 0) and crz(-1 -> 1) inside rcontrol (telegate)."""
    a = hea_ansatz(n_each, layers, seed, measure=False)
    b = hea_ansatz(n_each, layers, seed + 1, measure=False)
    a.append({"name": "expose", "qubits": [n_each - 1], "qpus": [ids[1]]})
    b.append({"name": "rcontrol", "qpus": [ids[0]], "instructions": [
        {"name": "cx", "qubits": [-1, 0]},
        {"name": "crz", "qubits": [-1, 1], "params": [0.7]},
    ]})
    a += measure_all(n_each)
    b += measure_all(n_each)
    return [
        {"id": ids[0], "num_qubits": n_each, "num_clbits": n_each, "instructions": a},
        {"id": ids[1], "num_qubits": n_each, "num_clbits": n_each, "instructions": b},
    ]


def config5_groups(n_groups=64, n_each=11, layers=2, shots=1024, seed=5, kinds=("teledata", "telegate"), **cfg):
    """BASELINE config 5: `n_groups` independent quantum-communication groups, each = two n_each-qubit vQPUs running a
    hardware-efficient ansatz (own random angles per group) joined by teledata or telegate (2 communication qubits):
    a (2 * n_each + 2)-qubit joint register per group (24 qubits at n_each = 11).  -> list of messages, one per group
    (each a list of two wire-format tasks, what the QC executor hands the engine in one round)."""
    msgs = []
    for g in range(n_groups):
        kind = kinds[g % len(kinds)]
        maker = hea_teledata_pair if kind == "teledata" else hea_telegate_pair
        tasks = maker(n_each, layers, seed=seed + 2 * g, ids=(f"g{g}A", f"g{g}B"))
        msgs.append([task(t["instructions"], t["num_qubits"], shots=shots, seed=seed + g, task_id=t["id"], is_dynamic=True, **cfg)
                     for t in tasks])
    return msgs


def n_gates(instructions):
    """number of original circuit gates (everything except measure/barrier), the gates/s numerator."""
    return sum(1 for i in instructions if i["name"] not in ("measure", "barrier", "save_state"))
