"""CPU: the engine's gate table (csrc/gates.cpp, read through cb200_gate_matrix) against an INDEPENDENT source.

tests/test_abi.py compares gates.cpp with the oracle's table -- both typed by the same hand, so a shared wrong
definition would pass.  Here every gate is rebuilt from its defining identity instead: Pauli algebra, matrix
exponentials of Pauli strings (scipy.linalg.expm), products of already-checked gates, and the controlled-gate
projector formula.  Names and conventions are Qiskit's (the reference maps its IR names 1:1 onto Qiskit
instructions, /root/reference/cunqa/qiskit_deps/transpiler.py; vocabulary /root/reference/src/utils/constants.hpp.in:271-465).
Matrix index bit m <-> qubit m of the instruction (little endian): kron(A_on_q1, B_on_q0).
"""
import numpy as np
import pytest
from scipy.linalg import expm

import cunqa_b200 as cb

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
P1 = np.array([[0, 0], [0, 1]], dtype=complex)
TOL = 1e-14


def kron(*ops):
    """kron(A_q(k-1), ..., A_q0): leftmost factor acts on the highest qubit"""
    out = np.eye(1, dtype=complex)
    for o in ops:
        out = np.kron(out, o)
    return out


def rot(pauli, theta):
    return expm(-0.5j * theta * pauli)


def controlled(u, nc):
    """controls are the first nc qubits (low index bits), targets the last: |1..1><1..1| (x) U + (1 - |1..1><1..1|) (x) I"""
    proj = kron(*([P1] * nc)) if nc else np.eye(1)
    dim_c = 1 << nc
    return np.kron(u, proj) + np.kron(np.eye(u.shape[0]), np.eye(dim_c) - proj)


def gm(name, nq, params=()):
    return cb.gate_matrix(name, nq, list(params))


def close(a, b):
    return np.abs(a - b).max() < TOL


# ---- one-qubit gates ------------------------------------------------------------------------------
def test_paulis_and_clifford(lib_built):
    assert close(gm("x", 1), X) and close(gm("y", 1), Y) and close(gm("z", 1), Z) and close(gm("id", 1), I2)
    assert close(gm("y", 1), 1j * X @ Z)                      # Y = iXZ
    h = gm("h", 1)
    assert close(h, (X + Z) / np.sqrt(2)) and close(h @ Z @ h, X)
    s, t = gm("s", 1), gm("t", 1)
    assert close(s @ s, Z) and close(t @ t, s) and s[0, 0] == 1 and t[0, 0] == 1
    assert close(gm("sdg", 1), s.conj().T) and close(gm("tdg", 1), t.conj().T)
    sx = gm("sx", 1)
    assert close(sx @ sx, X) and close(sx, np.exp(0.25j * np.pi) * rot(X, np.pi / 2))
    assert close(gm("sxdg", 1), sx.conj().T)
    sy = gm("sy", 1)
    assert close(sy @ sy, Y) and close(gm("sydg", 1), sy.conj().T)
    for alias, base in [("sz", "s"), ("szdg", "sdg"), ("v", "sx"), ("vdg", "sxdg")]:
        assert close(gm(alias, 1), gm(base, 1))


@pytest.mark.parametrize("theta", [0.0, 0.37, -2.1, np.pi, 5.9])
def test_rotations_are_pauli_exponentials(lib_built, theta):
    assert close(gm("rx", 1, [theta]), rot(X, theta))
    assert close(gm("ry", 1, [theta]), rot(Y, theta))
    assert close(gm("rz", 1, [theta]), rot(Z, theta))
    assert close(gm("p", 1, [theta]), np.exp(0.5j * theta) * rot(Z, theta))
    assert close(gm("u1", 1, [theta]), gm("p", 1, [theta]))
    phi = 0.8 - theta
    assert close(gm("r", 1, [theta, phi]), rot(np.cos(phi) * X + np.sin(phi) * Y, theta))


@pytest.mark.parametrize("t,p,l", [(0.3, 1.1, -0.7), (2.9, -2.2, 0.4), (np.pi / 2, 0.0, np.pi)])
def test_u_family(lib_built, t, p, l):
    # U(theta, phi, lambda) = e^{i(phi+lambda)/2} RZ(phi) RY(theta) RZ(lambda)
    want = np.exp(0.5j * (p + l)) * rot(Z, p) @ rot(Y, t) @ rot(Z, l)
    assert close(gm("u3", 1, [t, p, l]), want)
    assert close(gm("u", 1, [t, p, l]), want)
    assert close(gm("u2", 1, [p, l]), np.exp(0.5j * (p + l)) * rot(Z, p) @ rot(Y, np.pi / 2) @ rot(Z, l))
    g = 0.45  # 4-parameter form (quantum_task.cpp:88-96, apply_mcu :413): an extra global phase on the target block
    assert close(gm("u", 1, [t, p, l, g]), np.exp(1j * g) * want)
    assert close(gm("cu", 2, [t, p, l, g]), controlled(np.exp(1j * g) * want, 1))


# ---- two-qubit gates ------------------------------------------------------------------------------
def test_two_qubit_fixed_gates(lib_built):
    assert close(gm("swap", 2), 0.5 * (kron(I2, I2) + kron(X, X) + kron(Y, Y) + kron(Z, Z)))
    assert close(gm("iswap", 2), expm(0.25j * np.pi * (kron(X, X) + kron(Y, Y))))
    # ECR = (IX - XY)/sqrt(2), Qiskit little-endian: the left tensor factor acts on qubit 1
    ecr = gm("ecr", 2)
    assert close(ecr, (kron(I2, X) - kron(X, Y)) / np.sqrt(2)) and close(ecr @ ecr, np.eye(4))
    cx01 = controlled(X, 1)                       # control q0, target q1
    swap = gm("swap", 2)
    cx10 = swap @ cx01 @ swap                     # control q1, target q0
    assert close(gm("dcx", 2), cx10 @ cx01)       # DCX q0,q1 = cx q0,q1 ; cx q1,q0
    assert close(swap, cx01 @ cx10 @ cx01)


@pytest.mark.parametrize("theta", [0.41, -1.7, 3.0])
def test_two_qubit_rotations(lib_built, theta):
    assert close(gm("rxx", 2, [theta]), rot(kron(X, X), theta))
    assert close(gm("ryy", 2, [theta]), rot(kron(Y, Y), theta))
    assert close(gm("rzz", 2, [theta]), rot(kron(Z, Z), theta))
    assert close(gm("rzx", 2, [theta]), rot(kron(X, Z), theta))      # Z on qubit 0, X on qubit 1
    # XX-YY / XX+YY: at beta = 0 the plain exponential; beta enters as an RZ(beta) conjugation
    m0 = expm(-0.25j * theta * (kron(X, X) - kron(Y, Y)))
    p0 = expm(-0.25j * theta * (kron(X, X) + kron(Y, Y)))
    assert close(gm("xxmyy", 2, [theta, 0.0]), m0) and close(gm("xxpyy", 2, [theta, 0.0]), p0)
    beta = 0.9
    rz1 = kron(rot(Z, beta), I2)                                      # RZ(beta) on qubit 1
    assert close(gm("xxmyy", 2, [theta, beta]), rz1 @ m0 @ rz1.conj().T)
    rz0m = kron(I2, rot(Z, -beta))                                    # RZ(-beta) on qubit 0
    assert close(gm("xxpyy", 2, [theta, beta]), rz0m @ p0 @ rz0m.conj().T)
    # the published matrices (Qiskit XXMinusYYGate / XXPlusYYGate docs)
    c, s_ = np.cos(theta / 2), np.sin(theta / 2)
    assert close(gm("xxmyy", 2, [theta, beta]), np.array([[c, 0, 0, -1j * s_ * np.exp(-1j * beta)], [0, 1, 0, 0], [0, 0, 1, 0],
                                                           [-1j * s_ * np.exp(1j * beta), 0, 0, c]]))
    assert close(gm("xxpyy", 2, [theta, beta]), np.array([[1, 0, 0, 0], [0, c, -1j * s_ * np.exp(-1j * beta), 0],
                                                           [0, -1j * s_ * np.exp(1j * beta), c, 0], [0, 0, 0, 1]]))


# ---- controlled families: built from the (checked) base gate with the projector formula ------------------
CONTROLLED = [("cx", "x", 0), ("cy", "y", 0), ("cz", "z", 0), ("ch", "h", 0), ("cs", "s", 0), ("csdg", "sdg", 0), ("ct", "t", 0),
              ("csx", "sx", 0), ("csxdg", "sxdg", 0), ("crx", "rx", 1), ("cry", "ry", 1), ("crz", "rz", 1), ("cp", "p", 1),
              ("cu1", "p", 1), ("cu2", "u2", 2), ("cu3", "u3", 3), ("cr", "r", 2), ("cecr", "ecr", 0), ("cswap", "swap", 0)]
MULTI = [("ccx", "x", 0, 2), ("ccy", "y", 0, 2), ("ccz", "z", 0, 2), ("mcx", "x", 0, 3), ("mcy", "y", 0, 2), ("mcz", "z", 0, 3),
         ("mch", "h", 0, 2), ("mcs", "s", 0, 2), ("mct", "t", 0, 2), ("mcsx", "sx", 0, 2), ("mcp", "p", 1, 3), ("mcu1", "p", 1, 2),
         ("mcrx", "rx", 1, 2), ("mcry", "ry", 1, 3), ("mcrz", "rz", 1, 2), ("mcu2", "u2", 2, 2), ("mcu3", "u3", 3, 2),
         ("mcu", "u3", 3, 2), ("mcr", "r", 2, 2), ("mcswap", "swap", 0, 2)]


@pytest.mark.parametrize("name,base,npar", CONTROLLED)
def test_singly_controlled(lib_built, name, base, npar):
    params = list(np.random.default_rng(len(name) * 7 + npar).uniform(-3, 3, npar))
    nt = 2 if base in ("swap", "ecr") else 1
    assert close(gm(name, nt + 1, params), controlled(gm(base, nt, params), 1))


@pytest.mark.parametrize("name,base,npar,nc", MULTI)
def test_multi_controlled(lib_built, name, base, npar, nc):
    params = list(np.random.default_rng(len(name) * 11 + nc).uniform(-3, 3, npar))
    nt = 2 if base == "swap" else 1
    assert close(gm(name, nt + nc, params), controlled(gm(base, nt, params), nc))


def test_cz_and_cp_are_symmetric_and_ccx_is_toffoli(lib_built):
    swap = gm("swap", 2)
    assert close(swap @ gm("cz", 2) @ swap, gm("cz", 2))
    assert close(swap @ gm("cp", 2, [0.7]) @ swap, gm("cp", 2, [0.7]))
    tof = np.eye(8)
    tof[[3, 7]] = tof[[7, 3]]          # controls q0,q1 = 1 flips q2: |011> <-> |111>
    assert close(gm("ccx", 3), tof)
    h2 = kron(gm("h", 1), I2, I2)
    assert close(h2 @ gm("ccz", 3) @ h2, gm("ccx", 3))
