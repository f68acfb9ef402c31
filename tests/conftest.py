import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib_built():
    """build libcunqa_b200.so (nvcc cross-compiles for sm_100a without a GPU) and the C oracle"""
    from cunqa_b200 import _build
    from oracle import oracle_np
    _build.build()
    oracle_np.build_c_oracle()
    return True


def random_circuit(n, n_gates, seed, max_unitary=3):
    """mixed wire-format circuit over the engine's whole vocabulary"""
    import numpy as np
    from cunqa_b200 import circuits as C
    rng = np.random.default_rng(seed)
    names1 = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sxdg", "id"]
    ins = []
    for _ in range(n_gates):
        r = int(rng.integers(0, 10))
        q = list(map(int, rng.permutation(n)[:min(n, 5)]))
        f = lambda k=1: list(map(float, rng.uniform(-3.2, 3.2, k)))
        if r == 0 or n == 1:
            ins.append({"name": names1[int(rng.integers(len(names1)))], "qubits": q[:1]})
        elif r == 1:
            ins.append({"name": ["rx", "ry", "rz", "p", "u1"][int(rng.integers(5))], "qubits": q[:1], "params": f()})
        elif r == 2:
            ins.append({"name": ["cx", "cy", "cz", "swap", "ecr", "csx", "iswap", "dcx", "ch"][int(rng.integers(9))], "qubits": q[:2]})
        elif r == 3:
            ins.append({"name": ["cp", "crx", "cry", "crz", "rxx", "ryy", "rzz", "rzx", "cu1"][int(rng.integers(9))], "qubits": q[:2], "params": f()})
        elif r == 4 and n >= 3:
            ins.append({"name": ["ccx", "ccz", "cswap", "ccy"][int(rng.integers(4))], "qubits": q[:3]})
        elif r == 5:
            nm = ["u3", "u2", "r", "u"][int(rng.integers(4))]
            ins.append({"name": nm, "qubits": q[:1], "params": f({"u3": 3, "u2": 2, "r": 2, "u": 3}[nm])})
        elif r == 6 and n >= 4:
            nm = ["mcx", "mcp", "mcrz", "mcswap", "mcy", "mcz", "mcrx", "mcry", "mcsx", "mcu1"][int(rng.integers(10))]
            k = int(rng.integers(3, min(n, 5) + 1))
            ins.append({"name": nm, "qubits": q[:k], "params": f()})
        elif r == 7:
            k = int(rng.integers(1, min(n, max_unitary) + 1))
            ins.append(C.unitary_inst(C.haar_unitary(rng, 1 << k), q[:k]))
        elif r == 8:
            k = int(rng.integers(1, min(n, 4) + 1))
            d = np.exp(1j * rng.uniform(0, 6.28, 1 << k))
            ins.append({"name": "diagonal", "qubits": q[:k], "matrix": [[[float(z.real), float(z.imag)] for z in d]]})
        elif n >= 2:
            nm = ["cu3", "cu", "cu2", "cr"][int(rng.integers(4))]
            ins.append({"name": nm, "qubits": q[:2], "params": f({"cu3": 3, "cu": 4, "cu2": 2, "cr": 2}[nm])})
    return ins
