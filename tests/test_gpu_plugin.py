"""GPU (-m gpu): the C++ plugin classes LINKED and RUN end to end.

oracle/_ref/plugin_run (built by oracle/ref_boundary/build.sh where /root/reference exists; the binary travels to the GPU
box) links cunqa_b200/plugin/b200_plugin.cpp with the reference's own src/quantum_task.cpp + src/utils/json.cpp, an
in-process ClassicalChannel and libcunqa_b200.so, and drives it the way the reference's processes do:
  * SimulatorStrategy<SimpleBackend>::execute behind SimpleBackend::execute, fed by QuantumTask::update_circuit
    (simulator_strategy.hpp:9-17, simple_backend.hpp:82-85, qpu.cpp:33-63, aer_simple_simulator.cpp:8-22);
  * two CC vQPUs exchanging measured bits (aer_cc_simulator.cpp:17-33);
  * two QC vQPUs + the family's executor (aer_qc_simulator.cpp:22-32, aer_executor.cpp:41-85).
Results are compared with the CPU oracle.
"""
import json
import os
import subprocess

import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "plugin_run")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run_plugin(mode, doc, tmp_path, timeout=300):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/plugin_run not built (needs /root/reference at build time)")
    store = tmp_path / "store"
    (store / ".cunqa").mkdir(parents=True)
    env = dict(os.environ, STORE=str(store), SLURM_JOB_ID="4242", SLURM_TASK_PID="100")
    out = subprocess.run([BIN, mode], input=json.dumps(doc), capture_output=True, text=True, timeout=timeout, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return [json.loads(line) for line in out.stdout.splitlines() if line.strip()]


def test_simple_strategy_executes_the_reference_example_and_a_param_update(lib_built, tmp_path):
    # the canonical raw wire message, /root/reference/examples/cpp/simplest_example.cpp:8-28
    example = {"id": "circuit1", "config": {"shots": 10, "method": "automatic", "num_clbits": 2, "num_qubits": 2, "seed": 123123,
                                            "device": {"device_name": "CPU", "target_devices": []}},
               "instructions": [{"name": "h", "qubits": [0]}, {"name": "cx", "qubits": [0, 1]},
                                {"name": "measure", "qubits": [0], "clbits": [0]}, {"name": "measure", "qubits": [1], "clbits": [1]}],
               "sending_to": [], "is_dynamic": False}
    with open(os.path.join(GOLDEN, "hea6.json")) as f:
        rec = json.load(f)
    res = run_plugin("simple", {"messages": [example, rec["task"], rec["update"]]}, tmp_path)
    assert len(res) == 4
    assert all("ERROR" not in r for r in res[:3]), res
    _, want = O.run_static(example["instructions"], 2, shots=10, seed=123123, backend="numpy")
    assert res[0]["counts"] == want and set(res[0]["counts"]) <= {"00", "11"}
    cfg = rec["task"]["config"]
    _, want = O.run_static(rec["task"]["instructions"], cfg["num_qubits"], shots=cfg["shots"], seed=cfg["seed"], backend="numpy")
    assert res[1]["counts"] == want
    # {"params": ..} is applied by the REFERENCE's QuantumTask::update_params_ (quantum_task.cpp:39-108) inside the harness
    upd = cb.update_task(rec["task"], rec["update"])
    _, want = O.run_static(upd["instructions"], cfg["num_qubits"], shots=rec["update"]["shots"], seed=cfg["seed"], backend="numpy")
    assert res[2]["counts"] == want
    assert "noise" in res[3]["ERROR"]          # a backend configured with a noise model is refused, not silently ideal


def test_cc_strategies_exchange_measured_bits(lib_built, tmp_path):
    shots = 40
    a = [{"name": "h", "qubits": [0]}, {"name": "measure", "qubits": [0], "clbits": [0]},
         {"name": "send", "clbits": [0], "qpus": ["4242_102"]}]
    b = [{"name": "recv", "clbits": [0], "qpus": ["4242_101"]},
         {"name": "cif", "clbits": [0], "operation": "and", "condition": 1, "instructions": [{"name": "x", "qubits": [0]}]},
         {"name": "measure", "qubits": [0], "clbits": [1]}]
    ta = C.task(a, 1, 1, shots=shots, seed=5, task_id="A", is_dynamic=True, sending_to=["4242_102"])
    tb = C.task(b, 1, 2, shots=shots, seed=6, task_id="B", is_dynamic=True)
    ra, rb = run_plugin("cc", {"tasks": [ta, tb]}, tmp_path)
    assert "ERROR" not in ra and "ERROR" not in rb, (ra, rb)
    assert sum(ra["counts"].values()) == shots and sum(rb["counts"].values()) == shots
    assert set(rb["counts"]) <= {"00", "11"}                        # the received bit drives the feed-forward X
    assert rb["counts"].get("11", 0) == ra["counts"].get("1", 0)     # shot by shot the bit A measured
    want = O.run_dynamic([{"id": "A", "num_qubits": 1, "num_clbits": 1, "instructions": a[:2]}], shots, seed=5)["A"]
    assert ra["counts"] == want


def test_qc_strategies_and_executor_round(lib_built, tmp_path):
    shots = 300
    tasks = C.hea_telegate_pair(3, layers=1, seed=4)
    msgs = [C.task(t["instructions"], t["num_qubits"], shots=shots, seed=33, task_id=t["id"], is_dynamic=True) for t in tasks]
    res = run_plugin("qc", {"tasks": msgs, "rounds": 2}, tmp_path)
    assert len(res) == 4
    want = O.run_dynamic(tasks, shots, seed=33, defer=True)
    for r in res:
        assert "ERROR" not in r["result"], r
        tid = tasks[r["vqpu"]]["id"]
        assert r["result"]["counts"] == want[tid]
        assert r["result"]["time_taken"] > 0
