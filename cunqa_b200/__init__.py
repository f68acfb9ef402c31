"""cunqa_b200 -- B200-native state-vector backend for CUNQA.

The product is `libcunqa_b200.so` (hand-written sm_100a CUDA behind a C ABI, include/cunqa_b200.h)
plus the C++ plugin classes in cunqa_b200/plugin/ that slot into the reference's
src/backends/simulators/ directory.  This Python package is only the ctypes view of that C ABI
used by the tests and bench.py, plus synthetic workload generators.
"""
import importlib

from . import circuits  # noqa: F401

_LAZY = ("StateVector", "ShardedStateVector", "Backend", "B200Error", "plan", "gate_matrix", "version", "device_count", "update_task", "qasm2_to_json")


def __getattr__(name):
    # the binding is loaded lazily so that `import cunqa_b200.circuits` works before the library is built
    if name in _LAZY:
        return getattr(importlib.import_module("cunqa_b200.binding"), name)
    raise AttributeError(name)
