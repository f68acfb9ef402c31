"""OpenQASM 2 front end (host only): cb200_qasm2_to_json replaces the reference's qasm2_to_json
(src/utils/helpers/qasm2_to_json.hpp:520-549) -- same output object, checked here against hand-built circuits
through the CPU oracle."""
import math

import numpy as np
import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O

GHZ = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[4];
creg c[4];
h q[0];
cx q[0],q[1];
cx q[1],q[2]; cx q[2],q[3];   // two statements on one line
barrier q;
measure q -> c;
"""


def test_ghz_text_equals_builder(lib_built):
    circ = cb.qasm2_to_json(GHZ)
    assert circ["num_qubits"] == 4 and circ["num_clbits"] == 4
    assert circ["quantum_registers"] == {"q": [0, 1, 2, 3]} and circ["classical_registers"] == {"c": [0, 1, 2, 3]}
    assert circ["instructions"] == C.ghz(4)


def test_registers_are_numbered_in_declaration_order(lib_built):
    circ = cb.qasm2_to_json("qreg a[2]; qreg b[3]; creg m[1]; creg k[2];\nx b[1]; cx a[1],b[2]; measure b[1] -> k[0]; measure a -> k;")
    assert circ["num_qubits"] == 5 and circ["num_clbits"] == 3
    assert circ["quantum_registers"] == {"a": [0, 1], "b": [2, 3, 4]} and circ["classical_registers"] == {"m": [0], "k": [1, 2]}
    assert circ["instructions"] == [
        {"name": "x", "qubits": [3]}, {"name": "cx", "qubits": [1, 4]},
        {"name": "measure", "qubits": [3], "clbits": [1]},
        {"name": "measure", "qubits": [0], "clbits": [1]}, {"name": "measure", "qubits": [1], "clbits": [2]}]


def test_angle_expressions_and_broadcast(lib_built):
    circ = cb.qasm2_to_json("qreg q[3];\nrz(pi/2) q[0]; u3(pi/4, -pi, 2*pi/3) q[1]; rx(-(pi - 1.5)/2 + 0.25) q[2]; h q; cp(3*pi/8) q[0],q[2]; U(0.1,0.2,0.3) q[0]; CX q[0],q[1];")
    ins = circ["instructions"]
    assert ins[0] == {"name": "rz", "qubits": [0], "params": [pytest.approx(math.pi / 2)]}
    assert ins[1]["params"] == [pytest.approx(math.pi / 4), pytest.approx(-math.pi), pytest.approx(2 * math.pi / 3)]
    assert ins[2]["params"] == [pytest.approx(-(math.pi - 1.5) / 2 + 0.25)]
    assert [i["qubits"] for i in ins[3:6]] == [[0], [1], [2]] and all(i["name"] == "h" for i in ins[3:6])
    assert ins[6] == {"name": "cp", "qubits": [0, 2], "params": [pytest.approx(3 * math.pi / 8)]}
    assert ins[7]["name"] == "u3" and ins[8] == {"name": "cx", "qubits": [0, 1]}
    # the converted circuit runs: same state as the hand-written instruction list
    want = [{"name": "rz", "qubits": [0], "params": [math.pi / 2]}, {"name": "u3", "qubits": [1], "params": [math.pi / 4, -math.pi, 2 * math.pi / 3]},
            {"name": "rx", "qubits": [2], "params": [-(math.pi - 1.5) / 2 + 0.25]}] + [{"name": "h", "qubits": [q]} for q in range(3)] + \
           [{"name": "cp", "qubits": [0, 2], "params": [3 * math.pi / 8]}, {"name": "u3", "qubits": [0], "params": [0.1, 0.2, 0.3]}, {"name": "cx", "qubits": [0, 1]}]
    a, _ = O.run_static(ins, 3, backend="numpy")
    b, _ = O.run_static(want, 3, backend="numpy")
    assert np.abs(a.amplitudes() - b.amplitudes()).max() < 1e-14


@pytest.mark.parametrize("text,msg", [
    ("qreg q[2]; foo q[0];", "unsupported instruction"), ("qreg q[2]; cx q[0];", ""), ("qreg q[2]; h r[0];", "unknown quantum register"),
    ("qreg q[2]; h q[5];", "out of range"), ("qreg q[2]; rz(pi/) q[0];", "cannot evaluate"),
    ("qreg q[2]; creg c[2]; if(c==1) x q[0];", "not supported"), ("qreg q[2]; creg c[1]; measure q -> c;", "sizes differ"),
])
def test_malformed_text_is_an_error(lib_built, text, msg):
    with pytest.raises(cb.B200Error) as e:
        cb.qasm2_to_json(text)
    assert msg in str(e.value)
