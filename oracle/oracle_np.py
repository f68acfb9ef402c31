"""oracle_np.py -- numpy twin of the CPU oracle + wire-JSON interpreter.  TEST INFRASTRUCTURE ONLY.

Restates, independently of the CUDA product, what the reference computes on its CPU path:

* gate vocabulary / argument order: the cases of ``execute_shot_``
  (/root/reference/src/backends/simulators/AER/aer_adapters/aer_simulator_adapter.cpp:183-752)
  and ``AER_BASIS_GATES`` (/root/reference/src/utils/constants.hpp.in:468-485); gate *definitions*
  are the standard Qiskit matrices for those names (the IR maps 1:1 onto Qiskit
  ``Instruction(name=...)``, /root/reference/cunqa/qiskit_deps/transpiler.py).
* static run = evolve once + sample all shots, measure->clbit mapping honoured
  (aer_simulator_adapter.cpp:806-842 + aer_helpers.hpp:143-184: clbit 0 is the rightmost char).
* dynamic run = per-shot interpreter with the round-robin task scheduler, CIF, SEND/RECV,
  teledata (QSEND/QRECV) and telegate (EXPOSE/RCONTROL) (aer_simulator_adapter.cpp:120-792, 845-935).

PARITY UNPINNED: the reference's tests contain no numeric fixtures for this path;
this twin is pinned by analytic known answers and by agreement with oracle/sv_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import cmath
import ctypes
import math
import os
import subprocess
from collections import deque

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------
# counter-based RNG shared by oracle and product (splitmix64 finaliser)
# ----------------------------------------------------------------------------------------------
def uniform(seed: int, counter: int) -> float:
    z = (seed + (counter + 1) * 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    z = z ^ (z >> 31)
    return (z >> 11) * (1.0 / 9007199254740992.0)


def meas_counter(shot: int, ordinal: int) -> int:
    """RNG counter of the `ordinal`-th measurement/reset draw of shot `shot` (dynamic path)."""
    return ((shot + 1) << 20) + ordinal


# ----------------------------------------------------------------------------------------------
# gate definitions (Qiskit conventions; little-endian: matrix index bit m <-> qubits[m])
# ----------------------------------------------------------------------------------------------
def _u3(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -cmath.exp(1j * lam) * s],
                     [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]], dtype=complex)


_SQ2 = 1.0 / math.sqrt(2.0)
_FIXED_1Q = {
    "id": np.eye(2, dtype=complex),
    "x": np.array([[0, 1], [1, 0]], dtype=complex),
    "y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "z": np.diag([1, -1]).astype(complex),
    "h": np.array([[1, 1], [1, -1]], dtype=complex) * _SQ2,
    "s": np.diag([1, 1j]).astype(complex),
    "sdg": np.diag([1, -1j]).astype(complex),
    "t": np.diag([1, cmath.exp(1j * math.pi / 4)]).astype(complex),
    "tdg": np.diag([1, cmath.exp(-1j * math.pi / 4)]).astype(complex),
    "sx": np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex) / 2,
    "sxdg": np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=complex) / 2,
}
_FIXED_1Q["sz"] = _FIXED_1Q["s"]
_FIXED_1Q["szdg"] = _FIXED_1Q["sdg"]
_FIXED_1Q["v"] = _FIXED_1Q["sx"]
_FIXED_1Q["vdg"] = _FIXED_1Q["sxdg"]
_FIXED_1Q["sy"] = np.array([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]], dtype=complex) / 2
_FIXED_1Q["sydg"] = _FIXED_1Q["sy"].conj().T

_FIXED_2Q = {
    "swap": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex),
    "iswap": np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex),
    "dcx": np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=complex),
    "ecr": np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]], dtype=complex) * _SQ2,
}


def _param_1q(name, p):
    if name == "rx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if name == "ry":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == "rz":
        return np.diag([cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])]).astype(complex)
    if name in ("p", "u1"):
        return np.diag([1, cmath.exp(1j * p[0])]).astype(complex)
    if name == "u2":
        return _u3(math.pi / 2, p[0], p[1])
    if name in ("u3", "u"):
        m = _u3(p[0], p[1], p[2])
        if len(p) > 3:
            m = m * cmath.exp(1j * p[3])
        return m
    if name == "r":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * cmath.exp(-1j * p[1]) * s],
                         [-1j * cmath.exp(1j * p[1]) * s, c]], dtype=complex)
    raise KeyError(name)


def _param_2q(name, p):
    c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
    if name == "rxx":
        return np.array([[c, 0, 0, -1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [-1j * s, 0, 0, c]], dtype=complex)
    if name == "ryy":
        return np.array([[c, 0, 0, 1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [1j * s, 0, 0, c]], dtype=complex)
    if name == "rzz":
        a, b = cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])
        return np.diag([a, b, b, a]).astype(complex)
    if name == "rzx":
        return np.array([[c, 0, -1j * s, 0], [0, c, 0, 1j * s], [-1j * s, 0, c, 0], [0, 1j * s, 0, c]], dtype=complex)
    if name == "xxmyy":
        e = cmath.exp(1j * p[1])
        return np.array([[c, 0, 0, -1j * s / e], [0, 1, 0, 0], [0, 0, 1, 0], [-1j * s * e, 0, 0, c]], dtype=complex)
    if name == "xxpyy":
        e = cmath.exp(1j * p[1])
        return np.array([[1, 0, 0, 0], [0, c, -1j * s / e, 0], [0, -1j * s * e, c, 0], [0, 0, 0, 1]], dtype=complex)
    raise KeyError(name)


# name -> (base gate name, n_targets); all leading qubits are controls (Qiskit order: controls first)
_CONTROLLED = {
    "cx": ("x", 1), "ccx": ("x", 1), "mcx": ("x", 1),
    "cy": ("y", 1), "ccy": ("y", 1), "mcy": ("y", 1),
    "cz": ("z", 1), "ccz": ("z", 1), "mcz": ("z", 1),
    "ch": ("h", 1), "mch": ("h", 1),
    "cs": ("s", 1), "mcs": ("s", 1), "csdg": ("sdg", 1), "ct": ("t", 1), "mct": ("t", 1),
    "csx": ("sx", 1), "mcsx": ("sx", 1), "csxdg": ("sxdg", 1),
    "crx": ("rx", 1), "mcrx": ("rx", 1), "cry": ("ry", 1), "mcry": ("ry", 1),
    "crz": ("rz", 1), "mcrz": ("rz", 1),
    "cp": ("p", 1), "mcp": ("p", 1), "cu1": ("p", 1), "mcu1": ("p", 1),
    "cu2": ("u2", 1), "mcu2": ("u2", 1), "cu3": ("u3", 1), "mcu3": ("u3", 1),
    "cu": ("u", 1), "mcu": ("u", 1), "cr": ("r", 1), "mcr": ("r", 1),
    "cswap": ("swap", 2), "mcswap": ("swap", 2), "cecr": ("ecr", 2),
}


def base_matrix(name: str, params) -> np.ndarray:
    if name in _FIXED_1Q:
        return _FIXED_1Q[name]
    if name in _FIXED_2Q:
        return _FIXED_2Q[name]
    if name in ("rx", "ry", "rz", "p", "u1", "u2", "u3", "u", "r"):
        return _param_1q(name, params)
    return _param_2q(name, params)


def gate_spec(name: str, qubits, params):
    """-> (matrix over the targets, target qubits, control qubits)"""
    if name in _CONTROLLED:
        base, nt = _CONTROLLED[name]
        m = base_matrix(base, params)
        return m, list(qubits[-nt:]), list(qubits[:-nt])
    m = base_matrix(name, params)
    return m, list(qubits), []


def parse_matrix(nested) -> np.ndarray:
    """wire `matrix` field of unitary/cunitary: one extra list level, rows of [re, im]."""
    a = np.asarray(nested[0], dtype=float)
    return a[..., 0] + 1j * a[..., 1]


# ----------------------------------------------------------------------------------------------
# state back-ends
# ----------------------------------------------------------------------------------------------
class NumpyState:
    """2^n complex128, qubit 0 = LSB.  Pure numpy; for n <~ 22."""

    def __init__(self, n):
        self.n = n
        self.v = np.zeros(1 << n, dtype=complex)
        self.v[0] = 1.0

    def _axes(self, qubits):
        return [self.n - 1 - q for q in qubits]

    def apply_matrix(self, qubits, mat, ctrls=()):
        n, k = self.n, len(qubits)
        t = self.v.reshape((2,) * n)
        idx = [slice(None)] * n
        for c in ctrls:
            idx[n - 1 - c] = 1
        sub = t[tuple(idx)]
        # axes of the target qubits inside `sub` (control axes were dropped)
        remaining = [a for a in range(n) if (n - 1 - a) not in ctrls]
        ax = [remaining.index(n - 1 - q) for q in qubits]
        m = np.asarray(mat, dtype=complex).reshape((2,) * (2 * k))
        # matrix index bit m <-> qubits[m]; reshape puts the MSB first, so reverse
        out_axes = list(range(k))[::-1]
        in_axes = list(range(k, 2 * k))[::-1]
        res = np.tensordot(m, sub, axes=(in_axes, ax))  # result axes: out bits (MSB first) + rest
        # move the produced axes back to their positions
        res = np.moveaxis(res, out_axes, ax)
        t[tuple(idx)] = res
        self.v = t.reshape(-1)

    def apply_diagonal(self, qubits, diag):
        i = np.arange(1 << self.n)
        j = np.zeros_like(i)
        for m, q in enumerate(qubits):
            j |= ((i >> q) & 1) << m
        self.v = self.v * np.asarray(diag, dtype=complex)[j]

    def apply_swap(self, q0, q1, ctrls=()):
        self.apply_matrix([q0, q1], _FIXED_2Q["swap"], ctrls)

    def global_phase(self, theta):
        self.v = self.v * cmath.exp(1j * theta)

    def norm(self):
        return float(np.sum(np.abs(self.v) ** 2))

    def prob1(self, q):
        i = np.arange(1 << self.n)
        return float(np.sum(np.abs(self.v[((i >> q) & 1) == 1]) ** 2))

    def collapse(self, q, outcome, p):
        i = np.arange(1 << self.n)
        keep = ((i >> q) & 1) == outcome
        self.v = np.where(keep, self.v / math.sqrt(p), 0.0)

    def measure(self, q, seed, counter):
        total = self.norm()
        p1 = self.prob1(q)
        p0 = total - p1
        r = uniform(seed, counter) * total
        out = 1 if r >= p0 else 0
        self.collapse(q, out, (p1 if out else p0) / total)
        return out

    def sample(self, shots, seed):
        cdf = np.cumsum(np.abs(self.v) ** 2)
        u = np.array([uniform(seed, s) for s in range(shots)]) * cdf[-1]
        idx = np.searchsorted(cdf, u, side="right")
        return np.minimum(idx, (1 << self.n) - 1).astype(np.uint64)

    def amplitudes(self):
        return self.v


_LIB = None


def build_c_oracle(force=False, native=False):
    """compile oracle/sv_oracle.c -> oracle/libsv_oracle.so (gcc -O3 -fopenmp).

    The default build targets x86-64-v3 so that the .so built in the dev container also runs on the
    GPU box's host CPU; native=True (bench.py's CPU-baseline legs) rebuilds with -march=native on
    the machine that times it."""
    so = os.path.join(_HERE, "libsv_oracle_native.so" if native else "libsv_oracle.so")
    src = os.path.join(_HERE, "sv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native" if native else "-march=x86-64-v3", "-fopenmp",
                               "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def c_oracle(native=False):
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle(force=native, native=native))
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.orc_init_zero.argtypes = [dp, ctypes.c_int]
        lib.orc_apply_matrix.argtypes = [dp, ctypes.c_int, ip, ctypes.c_int, ip, ctypes.c_int, dp]
        lib.orc_apply_diagonal.argtypes = [dp, ctypes.c_int, ip, ctypes.c_int, dp]
        lib.orc_apply_swap.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int]
        lib.orc_norm.argtypes = [dp, ctypes.c_int]
        lib.orc_norm.restype = ctypes.c_double
        lib.orc_prob1.argtypes = [dp, ctypes.c_int, ctypes.c_int]
        lib.orc_prob1.restype = ctypes.c_double
        lib.orc_collapse.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        lib.orc_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        lib.orc_uniform.restype = ctypes.c_double
        lib.orc_measure.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64]
        lib.orc_measure.restype = ctypes.c_int
        lib.orc_sample.argtypes = [dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                   ctypes.POINTER(ctypes.c_uint64)]
        lib.orc_num_threads.restype = ctypes.c_int
        lib.orc_set_num_threads.argtypes = [ctypes.c_int]
        lib.orc_set_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _ia(xs):
    return (ctypes.c_int * max(1, len(xs)))(*xs)


class COracleState:
    """Same interface as NumpyState, arithmetic in oracle/sv_oracle.c (OpenMP)."""

    def __init__(self, n):
        self.n = n
        self.lib = c_oracle()
        self.v = np.zeros(1 << n, dtype=complex)
        self.lib.orc_init_zero(self._p(), n)

    def _p(self):
        return self.v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def apply_matrix(self, qubits, mat, ctrls=()):
        m = np.ascontiguousarray(np.asarray(mat, dtype=complex))
        self.lib.orc_apply_matrix(self._p(), self.n, _ia(qubits), len(qubits), _ia(list(ctrls)), len(ctrls),
                                  m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    def apply_diagonal(self, qubits, diag):
        d = np.ascontiguousarray(np.asarray(diag, dtype=complex))
        self.lib.orc_apply_diagonal(self._p(), self.n, _ia(qubits), len(qubits),
                                    d.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    def apply_swap(self, q0, q1, ctrls=()):
        self.lib.orc_apply_swap(self._p(), self.n, q0, q1, _ia(list(ctrls)), len(ctrls))

    def global_phase(self, theta):
        self.apply_diagonal([0], [cmath.exp(1j * theta)] * 2)

    def norm(self):
        return self.lib.orc_norm(self._p(), self.n)

    def prob1(self, q):
        return self.lib.orc_prob1(self._p(), self.n, q)

    def collapse(self, q, outcome, p):
        self.lib.orc_collapse(self._p(), self.n, q, outcome, p)

    def measure(self, q, seed, counter):
        return self.lib.orc_measure(self._p(), self.n, q, seed, counter)

    def sample(self, shots, seed):
        out = np.zeros(shots, dtype=np.uint64)
        self.lib.orc_sample(self._p(), self.n, shots, seed, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
        return out

    def amplitudes(self):
        return self.v


def make_state(n, backend="auto"):
    if backend == "numpy" or (backend == "auto" and n <= 12):
        return NumpyState(n)
    return COracleState(n)


# ----------------------------------------------------------------------------------------------
# instruction application (one gate at a time, no fusion)
# ----------------------------------------------------------------------------------------------
_NOOPS = ("barrier", "save_state")


def apply_gate(state, inst, qmap=lambda q: q, reverse_unitary_qubits=False):
    """apply one non-classical instruction; qmap maps instruction qubit -> register qubit."""
    name = inst["name"]
    if name in _NOOPS or name == "id":
        return
    qubits = [qmap(q) for q in inst.get("qubits", [])]
    params = inst.get("params", [])
    if name == "gp":
        state.global_phase(params[0])
    elif name == "unitary":
        m = parse_matrix(inst["matrix"])
        if reverse_unitary_qubits:  # quirk Q1 of the reference's dynamic path (aer_simulator_adapter.cpp:467)
            qubits = qubits[::-1]
        state.apply_matrix(qubits, m)
    elif name == "cunitary":
        m = parse_matrix(inst["matrix"])
        k = int(round(math.log2(m.shape[0])))
        state.apply_matrix(qubits[-k:], m, qubits[:-k])
    elif name == "diagonal":
        d = np.asarray(inst["matrix"][0], dtype=float)
        state.apply_diagonal(qubits, d[:, 0] + 1j * d[:, 1])
    elif name == "swap":
        state.apply_swap(qubits[0], qubits[1])
    else:
        m, tg, ct = gate_spec(name, qubits, params)
        state.apply_matrix(tg, m, ct)


def bits_to_string(creg, n_clbits):
    """clbit 0 is the rightmost character (aer_helpers.hpp:175-179, aer_simulator_adapter.cpp:780-789)"""
    return "".join("1" if creg.get(c, 0) else "0" for c in range(n_clbits - 1, -1, -1))


def run_static(instructions, n_qubits, n_clbits=None, shots=0, seed=0, backend="auto"):
    """evolve once, then sample all shots.  -> (state, counts dict or None)"""
    st = make_state(n_qubits, backend)
    meas = []
    for inst in instructions:
        if inst["name"] == "measure":
            for q, c in zip(inst["qubits"], inst["clbits"]):
                meas.append((q, c))
        else:
            apply_gate(st, inst)
    counts = None
    if shots:
        samples = st.sample(shots, seed)
        counts = {}
        width = n_clbits if n_clbits is not None else n_qubits
        for s in samples:
            s = int(s)
            creg = {c: (s >> q) & 1 for q, c in meas}
            key = bits_to_string(creg, width)
            counts[key] = counts.get(key, 0) + 1
    return st, counts


_CIF_OPS = {
    "and": lambda a, b: a & b, "or": lambda a, b: a | b, "xor": lambda a, b: a ^ b,
    "andn": lambda a, b: a & (1 - b), "orn": lambda a, b: a | (1 - b), "xorn": lambda a, b: a ^ (1 - b),
}


class DeferAbort(Exception):
    """the circuit needs a real mid-circuit collapse: measurements cannot all be deferred"""


class _NeedMoreBits(Exception):
    pass


class _DeadBranch(Exception):
    pass


def run_dynamic(tasks, shots, seed=0, n_comm_qubits=None, backend="auto", channel=None,
                reverse_unitary_qubits=False, forced_bits=None, teledata_literal=True, terminal_sampling=True,
                defer=False, _weight=None, _final=None):
    """Per-shot interpreter of one or several co-simulated tasks (aer_simulator_adapter.cpp:120-792,845-935).

    tasks: list of dicts {"id", "num_qubits", "num_clbits", "instructions"}.  Tasks are stepped
    round-robin in list order.  -> {task_id: {bitstring: count}}
    forced_bits: optional list; if given, measurement outcomes are taken from it in order
    (deterministic test mode; the state is collapsed onto the forced outcome).
    terminal_sampling: once every unfinished task has only `measure` instructions left, the remaining
    measure+collapse rounds are replaced by ONE inverse-CDF draw from |a|^2 (same joint distribution;
    this is the engine's documented shortcut, restated here so that seeded runs stay comparable shot by
    shot).  False = the reference's literal sequence of apply_measure calls (:185-190).
    teledata_literal: the reference's QRECV applies X on the *sent qubit's* outcome and Z on the
    *comm qubit's* outcome (aer_simulator_adapter.cpp:644-658, same in every adapter), which is the
    reverse of its own documented protocol (docs/source/_static/teledata.png: X <- comm qubit,
    Z <- sent qubit) and does not teleport.  The oracle restates the reference, so the default (True) is
    the literal code; False follows the documented protocol (the engine's
    `b200_teledata_corrections: "protocol"`; quirk Q6).
    defer: the engine's deferred-measurement mode, restated: ONE state for all shots; `measure q -> c` freezes q
    and makes c the symbol ("sym", q); a cif on such a bit and the teledata / telegate corrections become the
    same gate with q as an extra control; all shots are drawn from the final state (u_s = U(seed, s), as on
    the static path).  Raises DeferAbort when a frozen qubit would be driven again, a live qubit reset, or a
    condition is not a single undecided bit tested for 1 (the engine then falls back to the per-shot mode).
    """
    multi = len(tasks) > 1
    if n_comm_qubits is None:
        n_comm_qubits = 2 if multi else 0
    if n_comm_qubits % 2:
        n_comm_qubits += 1
    n_total = sum(t["num_qubits"] for t in tasks) + n_comm_qubits
    counters = {t["id"]: {} for t in tasks}

    def is_sym(v):
        return isinstance(v, tuple)

    for shot in range(1 if defer else shots):
        st = make_state(n_total, backend)
        draw = [0]
        forced = deque(forced_bits) if forced_bits is not None else None
        qstate = ["live"] * n_total   # defer: live / frozen / frozen_reset
        touched = [False] * n_total

        def measure(q, and_reset=False):
            if defer:
                if qstate[q] == "frozen_reset" or (qstate[q] == "live" and not touched[q]):
                    return 0
                qstate[q] = "frozen_reset" if and_reset else "frozen"
                return ("sym", q)
            if forced is not None:
                if not forced:
                    raise _NeedMoreBits()
                out = forced.popleft()
                total = st.norm()
                p1 = st.prob1(q)
                p = (p1 if out else total - p1) / total
                if _weight is not None:
                    _weight[0] *= p
                if p < 1e-14:
                    raise _DeadBranch()
                st.collapse(q, out, p)
            else:
                out = st.measure(q, seed, meas_counter(shot, draw[0]))
                draw[0] += 1
            if and_reset and out:
                st.apply_matrix([q], _FIXED_1Q["x"])
            return out

        def sym_referenced(q):
            v = ("sym", q)
            queues = list(q_td.values()) + list(q_tg.values()) + list(q_cc.values())
            return any(x == v for x in creg.values()) or any(x == v for dq in queues for x in dq)

        def recycle(q):
            """defer mode: bring a qubit back to |0> without a collapse -- possible when nothing still refers to its
            recorded outcome and it is disentangled from the rest (pure reduced state): one single-qubit unitary"""
            if sym_referenced(q):
                raise DeferAbort()
            a = st.amplitudes().reshape(1 << (n_total - q - 1), 2, 1 << q)
            a0, a1 = a[:, 0, :].ravel(), a[:, 1, :].ravel()
            r00, r11, r01 = float(np.vdot(a0, a0).real), float(np.vdot(a1, a1).real), complex(np.vdot(a1, a0))
            tr = r00 + r11
            purity = (r00 * r00 + r11 * r11 + 2 * abs(r01) ** 2) / (tr * tr)
            if not purity > 1 - 1e-9:
                raise DeferAbort()
            if r00 >= r11:
                phi0 = math.sqrt(r00 / tr)
                phi1 = np.conj(r01) / (tr * phi0)
            else:
                phi1 = math.sqrt(r11 / tr)
                phi0 = r01 / (tr * phi1)
            st.apply_matrix([q], np.array([[np.conj(phi0), np.conj(phi1)], [-phi1, phi0]], dtype=complex))
            qstate[q] = "live"
            touched[q] = False

        def reset(q):
            if defer:
                if qstate[q] == "frozen":
                    qstate[q] = "frozen_reset"  # record kept, logically |0> from here on
                elif qstate[q] == "live" and touched[q]:
                    recycle(q)
                return
            measure(q, True)

        def admit(inst, qubits):
            """defer mode: may this instruction still run?  False = drop it (controlled on a qubit that was measured
            and reset, logically |0>); DeferAbort = it would disturb a recorded outcome"""
            if any(qstate[q] != "live" for q in qubits):
                name = inst["name"]
                if name == "diagonal":
                    kinds = ["d"] * len(qubits)
                elif name in ("unitary", "cunitary"):
                    m = parse_matrix(inst["matrix"])
                    k = int(round(math.log2(m.shape[0])))
                    diag = np.count_nonzero(m - np.diag(np.diagonal(m))) == 0
                    kinds = ["k"] * (len(qubits) - k) + [("d" if diag else "x")] * k
                elif name == "gp":
                    kinds = ["d"] * len(qubits)
                elif name == "swap":
                    kinds = ["x", "x"]
                else:
                    m, tg, ct = gate_spec(name, qubits, inst.get("params", []))
                    diag = np.count_nonzero(m - np.diag(np.diagonal(m))) == 0
                    kinds = ["k"] * len(ct) + [("d" if diag else "x")] * len(tg)
                skip = False
                for q, kind in zip(qubits, kinds):
                    if qstate[q] == "frozen" and kind == "x":
                        raise DeferAbort()
                    if qstate[q] == "frozen_reset":
                        if kind == "k":
                            skip = True      # controlled on |0>: never fires
                        else:
                            recycle(q)       # driven again: needs the physical |0> (or DeferAbort)
                if skip:
                    return False
            for q in qubits:
                touched[q] = True
            return True

        def gate(inst, qmap=lambda q: q, cond=1):
            """one unitary instruction under a condition that is 0, 1 or a symbol (defer mode)"""
            if inst["name"] in _NOOPS or inst["name"] == "id" or cond == 0:
                return
            if not defer:
                apply_gate(st, inst, qmap, reverse_unitary_qubits)
                return
            qubits = [qmap(q) for q in inst.get("qubits", [])]
            # active reset, "measure q -> c; if (c) x q": the qubit is |0> afterwards, its record stays
            if inst["name"] == "x" and len(qubits) == 1 and cond == ("sym", qubits[0]) and qstate[qubits[0]] == "frozen":
                qstate[qubits[0]] = "frozen_reset"
                return
            if not admit(inst, qubits):
                return
            if not is_sym(cond):
                apply_gate(st, inst, qmap, reverse_unitary_qubits)
                return
            lits = list(cond[1]) if cond[0] == "and" else [cond]
            ctrls = [l[1] for l in lits]
            for l in lits:  # a literal that fires on outcome 0: X-conjugate that control qubit
                if l[0] == "nsym":
                    st.apply_matrix([l[1]], _FIXED_1Q["x"])
            gate_controlled(inst, qubits, ctrls)
            for l in lits:
                if l[0] == "nsym":
                    st.apply_matrix([l[1]], _FIXED_1Q["x"])

        def gate_controlled(inst, qubits, ctrls):
            name = inst["name"]
            ne = len(ctrls)
            if name == "gp":
                d = np.ones(1 << ne, dtype=complex)
                d[-1] = np.exp(1j * inst["params"][0])
                st.apply_diagonal(ctrls, d)
            elif name == "unitary":
                qs = qubits[::-1] if reverse_unitary_qubits else qubits
                st.apply_matrix(qs, parse_matrix(inst["matrix"]), ctrls)
            elif name == "cunitary":
                m = parse_matrix(inst["matrix"])
                k = int(round(math.log2(m.shape[0])))
                st.apply_matrix(qubits[-k:], m, qubits[:-k] + ctrls)
            elif name == "diagonal":
                d = np.asarray(inst["matrix"][0], dtype=float)
                full = np.ones(len(d) << ne, dtype=complex)
                full[-len(d):] = d[:, 0] + 1j * d[:, 1]
                st.apply_diagonal(qubits + ctrls, full)
            elif name == "swap":
                st.apply_swap(qubits[0], qubits[1], ctrls)
            else:
                m, tg, ct = gate_spec(name, qubits, inst.get("params", []))
                st.apply_matrix(tg, m, list(ct) + ctrls)

        def make_cond(lits):
            """conjunction of literals -> 0 (contradiction), one literal, or ("and", (lits...))"""
            lits = sorted(set(lits))
            for a in lits:
                for b in lits:
                    if a[1] == b[1] and a[0] != b[0]:
                        return 0
            return lits[0] if len(lits) == 1 else ("and", tuple(lits))

        def fixed(name, qubits, cond=1):
            gate({"name": name, "qubits": list(qubits)}, cond=cond)

        T = {}
        nq = ncl = 0
        for t in tasks:
            T[t["id"]] = dict(id=t["id"], ncl=t["num_clbits"], zq=nq, zc=ncl, pc=0, ins=t["instructions"],
                              finished=False, b_td=False, b_tg=False, b_cc=False, cat=False)
            nq += t["num_qubits"]
            ncl += t["num_clbits"]
        pairs = [dict(q0=nq + i, q1=nq + i + 1, idle=True, sendr=None, recvr=None, proto=None, label=0)
                 for i in range(0, n_comm_qubits, 2)]
        creg = {}
        q_td, q_tg, q_cc = {}, {}, {}

        def gen_ent(npairs):
            idle = [i for i, p in enumerate(pairs) if p["idle"]]
            if defer:  # pairs are interchangeable: take never-used ones first (a used pair would need a real reset)
                idle.sort(key=lambda i: touched[pairs[i]["q0"]] or touched[pairs[i]["q1"]])
            idx = idle[:npairs]
            if len(idx) < npairs:
                return []
            for i in idx:
                pairs[i]["idle"] = False
                reset(pairs[i]["q1"])
                reset(pairs[i]["q0"])
                fixed("h", [pairs[i]["q0"]])
                fixed("cx", [pairs[i]["q0"], pairs[i]["q1"]])
            return idx

        def my_pairs(sendr, recvr, proto, npairs=0):
            out = []
            for i, p in enumerate(pairs):
                if npairs and len(out) == npairs:
                    break
                if (not p["idle"]) and p["sendr"] == sendr and p["recvr"] == recvr and p["proto"] == proto:
                    out.append(i)
            return out

        def step(t, inst, comm_idx=(), cond=1):
            name = inst["name"]

            def qmap(q):
                if q < 0:
                    for i in comm_idx:
                        if (not pairs[i]["idle"]) and pairs[i]["label"] == q:
                            return pairs[i]["q1"]
                    raise RuntimeError("unresolved remote qubit")
                return q + t["zq"]

            if is_sym(cond) and name in ("measure", "copy", "recv"):
                raise DeferAbort()  # classical state written under an undecided condition
            if name == "measure":
                creg[inst["clbits"][0] + t["zc"]] = measure(inst["qubits"][0] + t["zq"])
            elif name == "copy":
                for l, r in zip(inst["l_clbits"], inst["r_clbits"]):
                    creg[l + t["zc"]] = creg.get(r + t["zc"], 0)
            elif name == "reset":
                reset(inst["qubits"][0] + t["zq"])
            elif name == "send":
                if multi:
                    key = (t["id"], inst["qpus"][0])
                    for c in inst["clbits"]:
                        q_cc.setdefault(key, deque()).append(creg.get(c + t["zc"], 0))
                else:
                    for c in inst["clbits"]:
                        channel.send_measure(creg.get(c + t["zc"], 0), inst["qpus"][0])
            elif name == "recv":
                if multi:
                    key = (inst["qpus"][0], t["id"])
                    if q_cc.get(key):
                        for c in inst["clbits"]:
                            creg[c + t["zc"]] = q_cc[key].popleft()
                        t["b_cc"] = False
                    else:
                        t["b_cc"] = True
                else:
                    for c in inst["clbits"]:
                        creg[c + t["zc"]] = 1 if channel.recv_measure(inst["qpus"][0]) == 1 else 0
            elif name == "cif":
                want = 1 if inst.get("condition", 1) else 0
                bits = [creg.get(c + t["zc"], 0) for c in inst["clbits"]]
                if any(is_sym(b) for b in bits):
                    # one recorded outcome tested for 1 or for 0, or the AND of several bits; an enclosing undecided
                    # condition joins the conjunction.  Anything else (or / xor folds) needs real bits.
                    never = cond == 0
                    if len(bits) == 1:
                        lits = [bits[0] if want else ("nsym", bits[0][1])]
                    else:
                        if not (inst["operation"] == "and" and want):
                            raise DeferAbort()
                        never = never or any(b == 0 for b in bits)
                        lits = [b for b in bits if is_sym(b)]
                    if is_sym(cond):
                        lits += list(cond[1]) if cond[0] == "and" else [cond]
                    res = 0 if never else make_cond(lits)
                    if res != 0:
                        for sub in inst["instructions"]:
                            step(t, sub, (), res)
                    return
                acc = bits[0] if want else 1 - bits[0]
                for b in bits[1:]:
                    acc = _CIF_OPS[inst["operation"]](acc, b)
                res = acc if want else 1 - acc
                if want == res:
                    for sub in inst["instructions"]:
                        step(t, sub, (), cond)
            elif name == "qsend":
                idx = gen_ent(1)
                if not idx:
                    t["b_td"] = True
                    return
                t["b_td"] = False
                p = pairs[idx[0]]
                p["proto"] = "teledata"
                q = inst["qubits"][0] + t["zq"]
                fixed("cx", [q, p["q0"]])
                fixed("h", [q])
                # measure + "reset if 1" (:622-624): the qubit is known to be |1>, so the reset is an X.  (The X comes
                # before the comm qubit's measurement here and after it in the reference: they act on different qubits.)
                q_td.setdefault(t["id"], deque()).append(measure(q, True))
                q_td[t["id"]].append(measure(p["q0"]))
                T[inst["qpus"][0]]["b_td"] = False
                p["sendr"], p["recvr"] = t["id"], inst["qpus"][0]
            elif name == "qrecv":
                src = inst["qpus"][0]
                if not q_td.get(src):
                    t["b_td"] = True
                    return
                if t["b_td"]:
                    return
                m1 = q_td[src].popleft()  # outcome of the sent qubit
                m2 = q_td[src].popleft()  # outcome of the sender's comm qubit
                p = pairs[my_pairs(src, t["id"], "teledata", 1)[0]]
                mx, mz = (m1, m2) if teledata_literal else (m2, m1)
                fixed("x", [p["q1"]], mx)
                fixed("z", [p["q1"]], mz)
                fixed("swap", [p["q1"], inst["qubits"][0] + t["zq"]])
                p["idle"] = True
            elif name == "expose":
                if not t["cat"]:
                    idx = gen_ent(len(inst["qubits"]))
                    if not idx:
                        t["b_tg"] = True
                        return
                    for qid, i in enumerate(idx):
                        p = pairs[i]
                        p["proto"] = "telegate"
                        p["label"] = -(qid + 1)
                        fixed("cx", [inst["qubits"][qid] + t["zq"], p["q0"]])
                        q_tg.setdefault(t["id"], deque()).append(measure(p["q0"]))
                        t["cat"] = True
                        t["b_tg"] = True
                        T[inst["qpus"][0]]["b_tg"] = False
                        p["sendr"], p["recvr"] = t["id"], inst["qpus"][0]
                    return
                for i in range(len(inst["qubits"])):
                    fixed("z", [inst["qubits"][i] + t["zq"]], q_tg[inst["qpus"][0]].popleft())
                t["cat"] = False
                for i in my_pairs(t["id"], inst["qpus"][0], "telegate", len(inst["qubits"])):
                    pairs[i]["idle"] = True
            elif name == "rcontrol":
                src = inst["qpus"][0]
                if not q_tg.get(src):
                    t["b_tg"] = True
                    return
                if t["b_tg"]:
                    return
                idx = my_pairs(src, t["id"], "telegate")
                for i in idx:
                    fixed("x", [pairs[i]["q1"]], q_tg[src].popleft())
                for sub in inst["instructions"]:
                    step(t, sub, idx, cond)
                for i in idx:
                    fixed("h", [pairs[i]["q1"]])
                    q_tg.setdefault(t["id"], deque()).append(measure(pairs[i]["q1"]))
                T[src]["b_tg"] = False
                t["b_tg"] = False
            else:
                gate(inst, qmap, cond)

        def try_terminal():
            if not terminal_sampling or forced is not None or defer:
                return False
            pend = []
            for t in T.values():
                if t["finished"]:
                    continue
                if t["b_td"] or t["b_tg"] or t["b_cc"]:
                    return False
                for inst in t["ins"][t["pc"]:]:
                    if inst["name"] != "measure":
                        return False
                    pend.append((inst["qubits"][0] + t["zq"], inst["clbits"][0] + t["zc"]))
            if len(pend) < 2:
                return False
            amps = st.amplitudes()
            cdf = np.cumsum(np.abs(amps) ** 2)
            u = uniform(seed, meas_counter(shot, draw[0])) * cdf[-1]
            draw[0] += 1
            idx = min(int(np.searchsorted(cdf, u, side="right")), len(cdf) - 1)
            for q, c in pend:
                creg[c] = (idx >> q) & 1
            for t in T.values():
                t["pc"] = len(t["ins"])
                t["finished"] = True
            return True

        ended = False
        guard = 0
        while not ended:
            if try_terminal():
                break
            ended = True
            for t in T.values():
                if t["finished"]:
                    continue
                if t["b_td"] or t["b_tg"] or t["b_cc"]:
                    ended = False
                    # a blocked task re-tries its instruction (the reference re-enters apply_next_instr
                    # only once the flag has been cleared by the peer; RECV re-polls its queue)
                    if t["b_cc"]:
                        step(t, t["ins"][t["pc"]])
                        if not t["b_cc"]:
                            t["pc"] += 1
                            if t["pc"] == len(t["ins"]):
                                t["finished"] = True
                    continue
                if t["pc"] == len(t["ins"]):
                    t["finished"] = True
                    continue
                step(t, t["ins"][t["pc"]])
                if not (t["b_td"] or t["b_tg"] or t["b_cc"]):
                    t["pc"] += 1
                if t["pc"] != len(t["ins"]):
                    ended = False
                else:
                    t["finished"] = True
            guard += 1
            if guard > 1_000_000:
                raise RuntimeError("dynamic interpreter deadlock")

        if defer:
            # every shot is one draw from the final state; a symbolic bit reads its qubit from the drawn index
            def keys_of(index):
                out = {}
                for t in T.values():
                    local = {}
                    for c, v in creg.items():
                        if t["zc"] <= c < t["zc"] + t["ncl"]:
                            local[c - t["zc"]] = ((index >> v[1]) & 1) if is_sym(v) else v
                    out[t["id"]] = bits_to_string(local, t["ncl"])
                return out

            if _final is not None:  # exact joint distribution instead of samples
                probs = np.abs(st.amplitudes()) ** 2
                for index in np.nonzero(probs > 0)[0]:
                    key = tuple(sorted(keys_of(int(index)).items()))
                    _final[key] = _final.get(key, 0.0) + float(probs[index])
                return counters
            for s in st.sample(shots, seed):
                for tid, key in keys_of(int(s)).items():
                    counters[tid][key] = counters[tid].get(key, 0) + 1
            return counters
        for t in T.values():
            local = {c - t["zc"]: v for c, v in creg.items() if t["zc"] <= c < t["zc"] + t["ncl"]}
            key = bits_to_string(local, t["ncl"])
            counters[t["id"]][key] = counters[t["id"]].get(key, 0) + 1
    return counters


def run_dynamic_auto(tasks, shots, seed=0, **kw):
    """the engine's default policy: defer every measurement when the circuit allows it, else shot by shot"""
    try:
        return run_dynamic(tasks, shots, seed, defer=True, **kw), "deferred"
    except DeferAbort:
        return run_dynamic(tasks, shots, seed, **kw), "per-shot"


def deferred_distribution(tasks, **kw):
    """exact joint distribution {((task id, bitstring), ...): p} of the deferred-measurement run"""
    final = {}
    run_dynamic(tasks, 1, 0, defer=True, _final=final, **kw)
    return final


def literal_distribution(tasks, **kw):
    """exact joint distribution of the reference's literal per-shot semantics, by enumerating every measurement
    branch (depth-first over forced outcomes; branches of probability 0 are cut).  Small circuits only."""
    dist = {}
    stack = [[]]
    while stack:
        prefix = stack.pop()
        weight = [1.0]
        try:
            res = run_dynamic(tasks, 1, 0, forced_bits=prefix, _weight=weight, terminal_sampling=False, **kw)
        except _NeedMoreBits:
            stack.append(prefix + [0])
            stack.append(prefix + [1])
            continue
        except _DeadBranch:
            continue
        key = tuple(sorted((tid, next(iter(cnt))) for tid, cnt in res.items()))
        dist[key] = dist.get(key, 0.0) + weight[0]
    return dist


def fuse_blocks(instructions, kmax=4):
    """Greedy fusion of consecutive gates into dense blocks of <= kmax qubits (the idea of Aer's CPU fusion
    pass: fewer sweeps over the state at the price of larger matrices).  Used only by the CPU-baseline legs of
    bench.py so that the baseline is not handicapped by running gate by gate.  -> list of (qubits, matrix)"""
    blocks, cur_q, cur_ins = [], [], []

    def close():
        if not cur_ins:
            return
        k = len(cur_q)
        m = np.eye(1 << k, dtype=complex)
        for col in range(1 << k):
            st = NumpyState(k)
            st.v[:] = 0
            st.v[col] = 1
            for inst in cur_ins:
                apply_gate(st, inst, qmap=lambda q: cur_q.index(q))
            m[:, col] = st.v
        blocks.append((list(cur_q), m))

    for inst in instructions:
        if inst["name"] in ("measure", "barrier", "save_state", "id"):
            continue
        qs = list(inst["qubits"])
        union = cur_q + [q for q in qs if q not in cur_q]
        if len(union) > kmax and cur_ins:
            close()
            cur_q, cur_ins = [], []
            union = qs
        cur_q = union
        cur_ins.append(inst)
    close()
    return blocks


# ----------------------------------------------------------------------------------------------
# analytic known answers
# ----------------------------------------------------------------------------------------------
def qft_amplitudes(n, x, idx):
    """amplitudes of QFT|x> at basis indices idx (with the final bit-reversal swaps applied)."""
    idx = np.asarray(idx, dtype=np.uint64).astype(object)
    dim = 1 << n
    ph = np.array([((int(i) * x) % dim) / dim for i in idx], dtype=float)
    return np.exp(2j * np.pi * ph) / math.sqrt(dim)


def tvd(counts_a, counts_b):
    keys = set(counts_a) | set(counts_b)
    na, nb = sum(counts_a.values()), sum(counts_b.values())
    return 0.5 * sum(abs(counts_a.get(k, 0) / na - counts_b.get(k, 0) / nb) for k in keys)
