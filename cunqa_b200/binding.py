"""ctypes binding of libcunqa_b200.so -- the same C ABI the C++ plugin calls (include/cunqa_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import, and
every compute entry point raises `B200Error` when no sm_100-class GPU is usable.
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcunqa_b200.so")


class B200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"cunqa_b200 error {status}: {message}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m cunqa_b200._build` "
            "(cunqa_b200 has no pure-Python or CPU fallback)")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)
_c_u64_p = ctypes.POINTER(ctypes.c_uint64)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)
_c_i8_p = ctypes.POINTER(ctypes.c_int8)
_vp = ctypes.c_void_p

SEND_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p)
RECV_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p)

# every symbol include/cunqa_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cb200_version": (ctypes.c_char_p, []),
    "cb200_last_error": (ctypes.c_char_p, []),
    "cb200_device_count": (ctypes.c_int, []),
    "cb200_free": (None, [_vp]),
    "cb200_state_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "cb200_state_destroy": (ctypes.c_int, [_vp]),
    "cb200_state_reset": (ctypes.c_int, [_vp]),
    "cb200_state_num_qubits": (ctypes.c_int, [_vp]),
    "cb200_state_batch": (ctypes.c_int, [_vp]),
    "cb200_apply_matrix": (ctypes.c_int, [_vp, _c_int_p, ctypes.c_int, _c_int_p, ctypes.c_int, _c_dbl_p]),
    "cb200_apply_diagonal": (ctypes.c_int, [_vp, _c_int_p, ctypes.c_int, _c_dbl_p]),
    "cb200_apply_gate": (ctypes.c_int, [_vp, ctypes.c_char_p, _c_int_p, ctypes.c_int, _c_dbl_p, ctypes.c_int]),
    "cb200_apply_gate_masked": (ctypes.c_int, [_vp, ctypes.c_char_p, _c_int_p, ctypes.c_int, _c_dbl_p, ctypes.c_int, _c_u8_p]),
    "cb200_flush": (ctypes.c_int, [_vp]),
    "cb200_sync": (ctypes.c_int, [_vp]),
    "cb200_norm": (ctypes.c_int, [_vp, _c_dbl_p]),
    "cb200_prob1": (ctypes.c_int, [_vp, ctypes.c_int, _c_dbl_p]),
    "cb200_probabilities": (ctypes.c_int, [_vp, ctypes.c_int, _c_int_p, ctypes.c_int, _c_dbl_p]),
    "cb200_measure": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint64, _c_u64_p, _c_i8_p, _c_int_p]),
    "cb200_reset_qubit": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint64, _c_u64_p, _c_i8_p]),
    "cb200_sample": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _c_u64_p]),
    "cb200_get_amplitudes": (ctypes.c_int, [_vp, ctypes.c_int, _c_u64_p, ctypes.c_uint64, _c_dbl_p]),
    "cb200_set_amplitudes": (ctypes.c_int, [_vp, ctypes.c_int, _c_dbl_p]),
    "cb200_state_stream": (_vp, [_vp]),
    "cb200_backend_stream": (_vp, [_vp]),
    "cb200_stats": (ctypes.c_int, [_vp, _c_u64_p]),
    "cb200_plan_json": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cb200_update_task_json": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cb200_gate_matrix": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, _c_dbl_p, ctypes.c_int, _c_dbl_p]),
    "cb200_backend_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    "cb200_backend_create_multi": (ctypes.c_int, [ctypes.POINTER(_vp), _c_int_p, ctypes.c_int]),
    "cb200_state_create_sharded": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, _c_int_p, ctypes.c_int]),
    "cb200_backend_destroy": (ctypes.c_int, [_vp]),
    "cb200_backend_set_channel": (ctypes.c_int, [_vp, SEND_FN, RECV_FN, _vp]),
    "cb200_backend_execute": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cb200_qasm2_to_json": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cb200_backend_execute_many": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.POINTER(_vp)]),
    "cb200_backend_batch_info": (ctypes.c_int, [_vp, _c_u64_p]),
    "cb200_backend_stats": (ctypes.c_int, [_vp, _c_u64_p]),
    "cb200_backend_last_mode": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int)]),
    "cb200_dist_unique_id": (ctypes.c_int, [_c_u8_p]),
    "cb200_state_create_dist": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_u8_p]),
    "cb200_dist_export_handle": (ctypes.c_int, [_vp, _c_u8_p]),
    "cb200_dist_import_handles": (ctypes.c_int, [_vp, _c_u8_p]),
    "cb200_dist_stats": (ctypes.c_int, [_vp, _c_u64_p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


STAT_KEYS = ("passes", "gates", "fused_ops", "bytes", "launches", "relabelled_swaps", "h2d_bytes", "d2h_bytes")


def _check(status):
    if status != 0:
        raise B200Error(status, lib.cb200_last_error().decode())


def _ints(xs):
    xs = [int(x) for x in xs]
    return (ctypes.c_int * max(1, len(xs)))(*xs), len(xs)


def _dbls(xs):
    xs = [float(x) for x in xs]
    return (ctypes.c_double * max(1, len(xs)))(*xs), len(xs)


def _take_string(ptr):
    try:
        return ctypes.string_at(ptr).decode()
    finally:
        lib.cb200_free(ptr)


def version():
    return lib.cb200_version().decode()


def device_count():
    return lib.cb200_device_count()


def gate_matrix(name, nq, params=()):
    """matrix of a named gate as the engine builds it (controls expanded); host only."""
    p, np_ = _dbls(params)
    out = np.zeros((1 << nq, 1 << nq), dtype=complex)
    _check(lib.cb200_gate_matrix(name.encode(), nq, p, np_, out.ctypes.data_as(_c_dbl_p)))
    return out


def plan(task, options=None):
    """fused ops + tile passes of a wire-format task (dict or str); host only, no GPU needed."""
    text = task if isinstance(task, str) else json.dumps(task)
    opts = json.dumps(options).encode() if options else None
    out = _vp()
    _check(lib.cb200_plan_json(text.encode(), opts, ctypes.byref(out)))
    return json.loads(_take_string(out))


def update_task(task, update):
    """apply a {"params":[...],"shots":N} message to a task the way the backend does; host only."""
    out = _vp()
    _check(lib.cb200_update_task_json(json.dumps(task).encode(), json.dumps(update).encode(), ctypes.byref(out)))
    return json.loads(_take_string(out))


def qasm2_to_json(qasm):
    """OpenQASM 2 text -> CUNQA circuit dict (instructions, num_qubits, num_clbits, registers); host only."""
    out = _vp()
    _check(lib.cb200_qasm2_to_json(qasm.encode(), ctypes.byref(out)))
    return json.loads(_take_string(out))


class StateVector:
    """cb200_state: n qubits x batch independent states on one B200."""

    def __init__(self, n_qubits, precision="double", device=0, batch=1, devices=None):
        prec = {"double": 64, "single": 32, 64: 64, 32: 32}[precision]
        self._h = _vp()
        if devices is not None and len(devices) > 1:
            # ONE process, several GPUs: the state is sharded over `devices` inside the library
            d, nd = _ints(devices)
            _check(lib.cb200_state_create_sharded(ctypes.byref(self._h), n_qubits, prec, d, nd))
        else:
            _check(lib.cb200_state_create(ctypes.byref(self._h), n_qubits, prec, device if devices is None else devices[0], batch))
        self.n = n_qubits
        self.batch = batch
        self.precision = prec

    def close(self):
        if getattr(self, "_h", None):
            lib.cb200_state_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(lib.cb200_state_reset(self._h))

    def apply_matrix(self, qubits, matrix, ctrls=()):
        q, k = _ints(qubits)
        c, nc = _ints(ctrls)
        m = np.ascontiguousarray(np.asarray(matrix, dtype=complex))
        _check(lib.cb200_apply_matrix(self._h, q, k, c, nc, m.ctypes.data_as(_c_dbl_p)))

    def apply_diagonal(self, qubits, diag):
        q, k = _ints(qubits)
        d = np.ascontiguousarray(np.asarray(diag, dtype=complex))
        _check(lib.cb200_apply_diagonal(self._h, q, k, d.ctypes.data_as(_c_dbl_p)))

    def apply_gate(self, name, qubits, params=(), flags=None):
        q, nq = _ints(qubits)
        p, np_ = _dbls(params)
        if flags is None:
            _check(lib.cb200_apply_gate(self._h, name.encode(), q, nq, p, np_))
        else:
            f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
            assert f.size == self.batch
            _check(lib.cb200_apply_gate_masked(self._h, name.encode(), q, nq, p, np_, f.ctypes.data_as(_c_u8_p)))

    def apply_instruction(self, inst):
        """one wire-format instruction dict (gates only)"""
        name = inst["name"]
        if name == "unitary":
            a = np.asarray(inst["matrix"][0], dtype=float)
            self.apply_matrix(inst["qubits"], a[..., 0] + 1j * a[..., 1])
        elif name == "diagonal":
            a = np.asarray(inst["matrix"][0], dtype=float)
            self.apply_diagonal(inst["qubits"], a[:, 0] + 1j * a[:, 1])
        else:
            self.apply_gate(name, inst["qubits"], inst.get("params", ()))

    def run(self, instructions):
        for inst in instructions:
            if inst["name"] != "measure":
                self.apply_instruction(inst)
        return self

    def flush(self):
        _check(lib.cb200_flush(self._h))

    def sync(self):
        _check(lib.cb200_sync(self._h))

    def norm(self):
        out = np.zeros(self.batch)
        _check(lib.cb200_norm(self._h, out.ctypes.data_as(_c_dbl_p)))
        return out

    def prob1(self, qubit):
        out = np.zeros(self.batch)
        _check(lib.cb200_prob1(self._h, qubit, out.ctypes.data_as(_c_dbl_p)))
        return out

    def probabilities(self, qubits, b=0):
        q, k = _ints(qubits)
        out = np.zeros(1 << k)
        _check(lib.cb200_probabilities(self._h, b, q, k, out.ctypes.data_as(_c_dbl_p)))
        return out

    def measure(self, qubit, seed=0, counters=None, forced=None, reset=False):
        ctr = np.ascontiguousarray(np.zeros(self.batch, dtype=np.uint64) if counters is None
                                   else np.asarray(counters, dtype=np.uint64))
        fp = None
        if forced is not None:
            fo = np.ascontiguousarray(np.asarray(forced, dtype=np.int8))
            fp = fo.ctypes.data_as(_c_i8_p)
        if reset:
            _check(lib.cb200_reset_qubit(self._h, qubit, seed, ctr.ctypes.data_as(_c_u64_p), fp))
            return None
        out = np.zeros(self.batch, dtype=np.int32)
        _check(lib.cb200_measure(self._h, qubit, seed, ctr.ctypes.data_as(_c_u64_p), fp, out.ctypes.data_as(_c_int_p)))
        return out

    def sample(self, shots, seed=0, b=0):
        out = np.zeros(max(1, shots), dtype=np.uint64)
        _check(lib.cb200_sample(self._h, b, shots, seed, out.ctypes.data_as(_c_u64_p)))
        return out[:shots]

    def amplitudes(self, idx=None, b=0):
        if idx is None:
            idx = np.arange(1 << self.n, dtype=np.uint64)
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.uint64))
        out = np.zeros(idx.size, dtype=complex)
        _check(lib.cb200_get_amplitudes(self._h, b, idx.ctypes.data_as(_c_u64_p), idx.size, out.ctypes.data_as(_c_dbl_p)))
        return out

    def set_amplitudes(self, vec, b=0):
        v = np.ascontiguousarray(np.asarray(vec, dtype=complex))
        assert v.size == 1 << self.n
        _check(lib.cb200_set_amplitudes(self._h, b, v.ctypes.data_as(_c_dbl_p)))

    def stats(self):
        out = np.zeros(8, dtype=np.uint64)
        _check(lib.cb200_stats(self._h, out.ctypes.data_as(_c_u64_p)))
        return {k: int(v) for k, v in zip(STAT_KEYS, out)}

    def stream(self):
        """cudaStream_t (as int) all kernels of this state run on"""
        return int(lib.cb200_state_stream(self._h) or 0)

    def dist_stats(self):
        out = np.zeros(5, dtype=np.uint64)
        _check(lib.cb200_dist_stats(self._h, out.ctypes.data_as(_c_u64_p)))
        return {"rank": int(out[0]), "world": int(out[1]), "swaps": int(out[2]), "swap_bytes_sent": int(out[3]),
                "swap_ms": int(out[4]) / 1000.0}


class ShardedStateVector(StateVector):
    """One state sharded over the GPUs of a box: one process per GPU, top log2(world) qubits global.

    Rendezvous uses an initialised torch.distributed process group (NCCL or gloo): rank 0's NCCL id is
    broadcast, every rank's CUDA IPC handle is all-gathered, then the engine talks NCCL / peer memory itself.
    """

    def __init__(self, n_qubits_total, precision="double", device=0, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        prec = {"double": 64, "single": 32, 64: 64, 32: 32}[precision]
        dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            _check(lib.cb200_dist_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(dev)
        dist.broadcast(ident, 0, group=group)
        idbuf = (ctypes.c_uint8 * 128)(*ident.cpu().tolist())
        self._h = _vp()
        _check(lib.cb200_state_create_dist(ctypes.byref(self._h), n_qubits_total, prec, device, rank, world, idbuf))
        self.n, self.batch, self.precision = n_qubits_total, 1, prec
        self.rank, self.world = rank, world
        hbuf = (ctypes.c_uint8 * 64)()
        _check(lib.cb200_dist_export_handle(self._h, hbuf))
        mine = torch.tensor(list(hbuf), dtype=torch.uint8).to(dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        flat = (ctypes.c_uint8 * (64 * world))(*[int(v) for h in allh for v in h.cpu().tolist()])
        _check(lib.cb200_dist_import_handles(self._h, flat))

    def dist_stats(self):
        out = np.zeros(5, dtype=np.uint64)
        _check(lib.cb200_dist_stats(self._h, out.ctypes.data_as(_c_u64_p)))
        return {"rank": int(out[0]), "world": int(out[1]), "swaps": int(out[2]), "swap_bytes_sent": int(out[3]),
                "swap_ms": int(out[4]) / 1000.0}


class Backend:
    """cb200_backend: what the CUNQA plugin's execute() drives (wire JSON in, result JSON out)."""

    def __init__(self, device=0, devices=None):
        self._h = _vp()
        if devices is not None and len(devices) > 1:
            d, nd = _ints(devices)
            _check(lib.cb200_backend_create_multi(ctypes.byref(self._h), d, nd))
        else:
            _check(lib.cb200_backend_create(ctypes.byref(self._h), device if devices is None else devices[0]))
        self._cb = None

    def close(self):
        if getattr(self, "_h", None):
            lib.cb200_backend_destroy(self._h)
            self._h = None

    __del__ = close

    def set_channel(self, send, recv):
        """send(bit:int, target:str) / recv(origin:str)->int, cf. ClassicalChannel::send_measure/recv_measure"""
        s = SEND_FN(lambda user, bit, tgt: send(int(bit), tgt.decode()))
        r = RECV_FN(lambda user, org: int(recv(org.decode())))
        self._cb = (s, r)
        _check(lib.cb200_backend_set_channel(self._h, s, r, None))

    def execute(self, task, parse=True):
        """task: dict / list of dicts / JSON text.  -> result dict ({"ERROR":...} on failure); parse=False: the JSON text"""
        text = task if isinstance(task, (str, bytes)) else json.dumps(task)
        out = _vp()
        _check(lib.cb200_backend_execute(self._h, text if isinstance(text, bytes) else text.encode(), ctypes.byref(out)))
        return json.loads(_take_string(out)) if parse else _take_string(out)

    def execute_many(self, messages, parse=True):
        """independent messages (dicts / lists / JSON texts) as separate registers batched on the device -> list of results"""
        texts = [(m if isinstance(m, str) else json.dumps(m)).encode() for m in messages]
        arr = (ctypes.c_char_p * max(1, len(texts)))(*texts)
        outs = (_vp * max(1, len(texts)))()
        _check(lib.cb200_backend_execute_many(self._h, arr, len(texts), outs))
        res = [_take_string(outs[i]) for i in range(len(texts))]
        return [json.loads(r) for r in res] if parse else res

    def batch_info(self):
        out = np.zeros(3, dtype=np.uint64)
        _check(lib.cb200_backend_batch_info(self._h, out.ctypes.data_as(_c_u64_p)))
        return {"uniform_flushes": int(out[0]), "split_flushes": int(out[1]), "solo": int(out[2])}

    def stats(self):
        out = np.zeros(8, dtype=np.uint64)
        _check(lib.cb200_backend_stats(self._h, out.ctypes.data_as(_c_u64_p)))
        return {k: int(v) for k, v in zip(STAT_KEYS, out)}

    def last_mode(self):
        """0 static, 1 dynamic with every measurement deferred, 2 dynamic on the batched per-shot interpreter"""
        out = ctypes.c_int(0)
        _check(lib.cb200_backend_last_mode(self._h, ctypes.byref(out)))
        return int(out.value)

    def stream(self):
        return int(lib.cb200_backend_stream(self._h) or 0)
