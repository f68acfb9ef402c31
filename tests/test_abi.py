"""CPU: the C-ABI library loads, exports every symbol include/cunqa_b200.h declares, its named
gates agree with the independent numpy definitions, and it fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import cunqa_b200 as cb
from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cunqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)) - {"cb200_send_fn", "cb200_recv_fn"})


def test_every_declared_symbol_is_exported(lib_built):
    from cunqa_b200 import binding
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(binding.lib, s), f"{s} declared in include/cunqa_b200.h but not exported"
    assert set(syms) == set(binding.SIGNATURES), "binding.SIGNATURES must list exactly the header's entry points"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "cunqa_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "at::" not in code and "std::" not in code


GATES = [("id", 1, 0), ("x", 1, 0), ("y", 1, 0), ("z", 1, 0), ("h", 1, 0), ("s", 1, 0), ("sdg", 1, 0), ("t", 1, 0),
         ("tdg", 1, 0), ("sx", 1, 0), ("sxdg", 1, 0), ("sy", 1, 0), ("sydg", 1, 0), ("v", 1, 0), ("vdg", 1, 0),
         ("rx", 1, 1), ("ry", 1, 1), ("rz", 1, 1), ("p", 1, 1), ("u1", 1, 1), ("u2", 1, 2), ("u3", 1, 3), ("u", 1, 3),
         ("u", 1, 4), ("r", 1, 2), ("swap", 2, 0), ("iswap", 2, 0), ("dcx", 2, 0), ("ecr", 2, 0), ("cx", 2, 0),
         ("cy", 2, 0), ("cz", 2, 0), ("ch", 2, 0), ("cs", 2, 0), ("csdg", 2, 0), ("ct", 2, 0), ("csx", 2, 0),
         ("csxdg", 2, 0), ("crx", 2, 1), ("cry", 2, 1), ("crz", 2, 1), ("cp", 2, 1), ("cu1", 2, 1), ("cu2", 2, 2),
         ("cu3", 2, 3), ("cu", 2, 4), ("cr", 2, 2), ("rxx", 2, 1), ("ryy", 2, 1), ("rzz", 2, 1), ("rzx", 2, 1),
         ("xxmyy", 2, 2), ("xxpyy", 2, 2), ("ccx", 3, 0), ("ccy", 3, 0), ("ccz", 3, 0), ("cswap", 3, 0),
         ("mcx", 4, 0), ("mcy", 3, 0), ("mcz", 4, 0), ("mcsx", 3, 0), ("mcp", 4, 1), ("mcu1", 3, 1), ("mcrx", 3, 1),
         ("mcry", 4, 1), ("mcrz", 3, 1), ("mcu2", 3, 2), ("mcu3", 3, 3), ("mcu", 3, 4), ("mcr", 3, 2), ("mcswap", 4, 0)]


@pytest.mark.parametrize("name,nq,npar", GATES)
def test_gate_matrix_matches_numpy_twin(lib_built, name, nq, npar):
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    params = list(rng.uniform(-3, 3, npar))
    got = cb.gate_matrix(name, nq, params)
    m, tg, ct = O.gate_spec(name, list(range(nq)), params)
    st_cols = []
    for c in range(1 << nq):  # column c of the full matrix = gate applied to basis state c
        st = O.NumpyState(nq)
        st.v[:] = 0
        st.v[c] = 1
        st.apply_matrix(tg, m, ct)
        st_cols.append(st.v.copy())
    ref = np.array(st_cols).T
    assert np.abs(got - ref).max() < 1e-14
    assert np.abs(got @ got.conj().T - np.eye(1 << nq)).max() < 1e-13


def test_fails_loudly_without_gpu(lib_built):
    if cb.device_count() > 0:
        pytest.skip("a GPU is present")
    from cunqa_b200.binding import B200Error
    with pytest.raises(B200Error, match="no CPU fallback"):
        cb.StateVector(3)
    be = cb.Backend(0)
    res = be.execute({"id": "t", "config": {"shots": 1, "num_qubits": 1, "num_clbits": 1}, "instructions": [], "is_dynamic": False})
    assert "ERROR" in res and "no CPU fallback" in res["ERROR"]
