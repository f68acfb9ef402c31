#!/usr/bin/env python
"""Generates tests/golden/*.json.  Run in the dev container (needs /root/reference); the fixtures are
committed so that the GPU box (no /root/reference) can use them.

  python tests/golden/gen_golden.py

What is pinned, and by which piece of the reference:
  * wire format: every circuit is built with the REFERENCE'S OWN Python front end
    (/root/reference/cunqa/circuit/core.py `CunqaCircuit`, `to_ir`), imported from a scratch copy with
    the two compiled modules stubbed exactly as the reference's CI does
    (/root/reference/.github/workflows/tests.yml:27-59);
  * parsing + parameter update: each task (and a {"params":..} update) is pushed through
    oracle/_ref/ref_parse, i.e. the reference's own QuantumTask::update_circuit/update_params_ and
    from_json_instructions_to_cunqainstructions compiled from /root/reference/src;
  * numbers: amplitudes / counts come from the CPU oracle (the reference keeps no numeric fixtures for
    this path -- parity unpinned), so these vectors pin the oracle against regressions
    and give the GPU tests a fixed target; analytic closed forms are stored alongside where they exist.
"""
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_frontend():
    tmp = tempfile.mkdtemp(prefix="cunqa_ref_")
    shutil.copytree(os.path.join(REF, "cunqa"), os.path.join(tmp, "cunqa"))
    with open(os.path.join(tmp, "cunqa", "qclient.py"), "w") as f:
        f.write("class QClient:\n    def connect(self, a):\n        pass\nclass FutureWrapper:\n    pass\n"
                "def write_on_file(a, b, c):\n    pass\ndef qasm2_to_json(a):\n    pass\ndef json_to_qasm2(a):\n    pass\n")
    with open(os.path.join(tmp, "cunqa", "counts_and_probs.py"), "w") as f:
        f.write("def marginalize_counts(counts, region_sizes, check_lenght=False):\n    pass\n"
                "def recombine_probs(probs, per_qubit, partial, num_qubits):\n    pass\n"
                "def counts_to_probs(counts, per_qubit=False, partial=None):\n    pass\n")
    src = open(os.path.join(tmp, "cunqa", "constants.py.in")).read()
    src = src.replace("CUNQA_USE_QISKIT_PY = @CUNQA_USE_QISKIT_PY@", "CUNQA_USE_QISKIT_PY = False")
    with open(os.path.join(tmp, "cunqa", "constants.py"), "w") as f:
        f.write(src)
    os.environ.setdefault("STORE", tmp)
    sys.path.insert(0, tmp)
    from cunqa.circuit import CunqaCircuit
    from cunqa.circuit.ir import to_ir
    return CunqaCircuit, to_ir


def wire_task(ir, shots, seed, **cfg):
    """the message cunqa/qjob.py:87-110 builds from an IR"""
    config = {"shots": shots, "method": "statevector", "avoid_parallelization": False,
              "num_clbits": ir["num_clbits"], "num_qubits": ir["num_qubits"], "seed": seed,
              "device": {"device_name": "GPU", "target_devices": [0]}}
    config.update(cfg)
    return {"id": ir["id"], "config": config, "instructions": ir["instructions"],
            "sending_to": ir["sending_to"], "is_dynamic": ir["is_dynamic"]}


def build_cases(CunqaCircuit, to_ir):
    cases = {}
    # 1. GHZ (BASELINE config 1, shrunk)
    n = 8
    c = CunqaCircuit(n, n, id="ghz8")
    c.h(0)
    for i in range(n - 1):
        c.cx(i, i + 1)
    c.measure(list(range(n)), list(range(n)))
    cases["ghz8"] = wire_task(to_ir(c), 500, 1234)
    # 2. QFT on |x>
    n, x = 7, 0b1011001
    c = CunqaCircuit(n, n, id="qft7")
    for q in range(n):
        if (x >> q) & 1:
            c.x(q)
    for j in range(n - 1, -1, -1):
        c.h(j)
        for k in range(j - 1, -1, -1):
            c.cp(math.pi / (1 << (j - k)), k, j)
    for i in range(n // 2):
        c.swap(i, n - 1 - i)
    c.measure(list(range(n)), list(range(n)))
    cases["qft7"] = wire_task(to_ir(c), 300, 7)
    # 3. a tour of the gate vocabulary through the reference's builder methods
    c = CunqaCircuit(5, 5, id="vocab5")
    c.h(0); c.x(1); c.y(2); c.z(3); c.s(4); c.sdg(0); c.t(1); c.tdg(2); c.sx(3); c.sxdg(4)
    c.rx(0.3, 0); c.ry(1.1, 1); c.rz(-0.7, 2); c.p(0.9, 3); c.u1(0.4, 4); c.u2(0.2, 0.5, 0); c.u3(0.1, 0.2, 0.3, 1)
    c.r(0.6, 0.8, 2); c.u(1.2, 0.4, 0.7, 3)
    c.cx(0, 1); c.cy(1, 2); c.cz(2, 3); c.csx(3, 4); c.swap(0, 4); c.ecr(1, 3)
    c.cp(0.5, 0, 2); c.cu1(0.3, 1, 3); c.crx(0.7, 2, 4); c.cry(0.9, 3, 0); c.crz(1.3, 4, 1)
    c.rxx(0.4, 0, 1); c.ryy(0.6, 1, 2); c.rzz(0.8, 2, 3); c.rzx(1.0, 3, 4)
    c.cu3(0.3, 0.6, 0.9, 0, 3); c.cu(0.2, 0.4, 0.6, 0.8, 1, 4)
    c.ccx(0, 1, 2); c.ccz(1, 2, 3); c.cswap(2, 3, 4)
    c.mcx(0, 1, 2, 3); c.mcz(1, 2, 3, 4); c.mcp(0.35, 0, 2, 4); c.mcrx(0.45, 1, 3, 0); c.mcry(0.55, 2, 4, 1)
    c.mcrz(0.65, 3, 0, 2); c.mcswap(0, 1, 2, 3)
    rng = np.random.default_rng(5)
    z = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    q, r = np.linalg.qr(z)
    u = q * (np.diagonal(r) / np.abs(np.diagonal(r)))
    c.unitary(u.tolist(), 2, 4)
    c.diagonal(list(np.exp(1j * rng.uniform(0, 6, 4))), 0, 3)
    c.measure(list(range(5)), list(range(5)))
    cases["vocab5"] = wire_task(to_ir(c), 400, 99)
    # 4. hardware-efficient ansatz (VQE parameter-update path)
    n = 6
    c = CunqaCircuit(n, n, id="hea6")
    th = np.linspace(0.1, 2.9, 2 * n * 2)
    p = 0
    for _ in range(2):
        for qb in range(n):
            c.ry(float(th[p]), qb); c.rz(float(th[p + 1]), qb); p += 2
        for qb in range(n - 1):
            c.cx(qb, qb + 1)
    c.measure(list(range(n)), list(range(n)))
    cases["hea6"] = wire_task(to_ir(c), 400, 5)
    cases["hea6"]["update"] = {"params": [float(v) for v in th[::-1]], "shots": 250}
    # 5. dynamic: mid-circuit measure + cif + reset
    c = CunqaCircuit(3, 3, id="dyn3")
    c.h(0); c.ry(0.9, 2)
    c.measure(0, 0)
    with c.cif(0) as sub:
        sub.x(1)
        sub.rx(1.3, 2)
    c.measure(1, 1); c.measure(2, 2)
    cases["dyn3"] = wire_task(to_ir(c), 300, 17)
    # 6 + 7. teledata / telegate pairs (quantum communication family)
    a = CunqaCircuit(2, 2, id="tdA"); b = CunqaCircuit(2, 2, id="tdB")
    a.ry(1.1, 0); a.h(1)
    a.qsend(0, "tdB")
    b.qrecv(0, "tdA"); b.cx(0, 1)
    a.measure([0, 1], [0, 1]); b.measure([0, 1], [0, 1])
    ia, ib = to_ir(a), to_ir(b)
    cases["teledata"] = {"tasks": [wire_task(ia, 400, 33), wire_task(ib, 400, 33)]}
    a = CunqaCircuit(2, 2, id="tgA"); b = CunqaCircuit(2, 2, id="tgB")
    a.h(0); a.ry(0.4, 1)
    with a.expose(0, b) as ([ctrl], rc):
        rc.cx(ctrl, 0)
        rc.crz(0.7, ctrl, 1)
    a.measure([0, 1], [0, 1]); b.measure([0, 1], [0, 1])
    ia, ib = to_ir(a), to_ir(b)
    cases["telegate"] = {"tasks": [wire_task(ia, 400, 33), wire_task(ib, 400, 33)]}
    return cases


def fix_remote_ids(tasks):
    """cunqa/qpu.py:235-239 rewrites "circuits" -> "qpus" before sending; do the same here"""
    def fix(insts):
        for inst in insts:
            if "circuits" in inst:
                inst["qpus"] = inst.pop("circuits")
            if "instructions" in inst:
                fix(inst["instructions"])
    for t in tasks:
        fix(t["instructions"])


def ref_parse(lines):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_parse")
    out = subprocess.run([exe], input="\n".join(json.dumps(l) for l in lines) + "\n", capture_output=True, text=True,
                         env={**os.environ, "STORE": "/tmp"})
    return [json.loads(l) for l in out.stdout.splitlines() if l.strip()]


def main():
    from oracle import oracle_np as O
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_boundary", "build.sh")])
    CunqaCircuit, to_ir = import_reference_frontend()
    cases = build_cases(CunqaCircuit, to_ir)
    out = {}
    for name, case in cases.items():
        rec = {"generated_by": "tests/golden/gen_golden.py (reference front end: cunqa.circuit.CunqaCircuit + to_ir)"}
        if "tasks" in case:
            fix_remote_ids(case["tasks"])
            rec["tasks"] = case["tasks"]
            rec["ref_parsed"] = [ref_parse([t])[0]["structured"] for t in case["tasks"]]
            tasks = [{"id": t["id"], "num_qubits": t["config"]["num_qubits"], "num_clbits": t["config"]["num_clbits"],
                      "instructions": t["instructions"]} for t in case["tasks"]]
            shots, seed = case["tasks"][0]["config"]["shots"], case["tasks"][0]["config"]["seed"]
            rec["oracle_id_counts"] = O.run_dynamic(tasks, shots, seed=seed)
        else:
            update = case.pop("update", None)
            fix_remote_ids([case])
            rec["task"] = case
            parsed = ref_parse([case] + ([update] if update else []))
            rec["ref_parsed"] = parsed[0]["structured"]
            n, ncl = case["config"]["num_qubits"], case["config"]["num_clbits"]
            shots, seed = case["config"]["shots"], case["config"]["seed"]
            if case["is_dynamic"]:
                rec["oracle_counts"] = O.run_dynamic([{"id": case["id"], "num_qubits": n, "num_clbits": ncl,
                                                       "instructions": case["instructions"]}], shots, seed=seed)[case["id"]]
            else:
                st, counts = O.run_static(case["instructions"], n, n_clbits=ncl, shots=shots, seed=seed, backend="numpy")
                rec["oracle_statevector"] = [[float(z.real), float(z.imag)] for z in st.amplitudes()]
                rec["oracle_counts"] = counts
            if update:
                rec["update"] = update
                rec["ref_parsed_after_update"] = parsed[1]["structured"]
                rec["ref_config_after_update"] = parsed[1]["config"]
        out[name] = rec
    # analytic anchors
    n, x = 7, 0b1011001
    out["qft7"]["analytic_statevector"] = [[float(z.real), float(z.imag)] for z in O.qft_amplitudes(n, x, np.arange(1 << n))]
    for name, rec in out.items():
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(HERE, name + ".json")), "bytes")


if __name__ == "__main__":
    main()
