"""CPU, world_size 2 (gloo): the sharded-state planning logic (cunqa_b200/csrc/dist_plan.cpp, obtained per
rank through the C ABI cb200_plan_json) replayed by two processes that each hold half of the state and
exchange slabs with torch.distributed send/recv exactly where the plan says "swap".

The gathered result must equal the single-process oracle run of the same circuit.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
X = np.array([[0, 1], [1, 0]], dtype=complex)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _apply_local(st, op):
    m = np.array([complex(*z) for z in op["m"]])
    if op["kind"] == "dense":
        d = 1 << len(op["q"])
        st.apply_matrix(op["q"], m.reshape(d, d), op["ctrl"])
    elif op["kind"] == "diag":
        st.apply_diagonal(op["q"], m)
    elif op["kind"] == "permx":
        st.apply_matrix(op["q"], X, op["ctrl"])
    else:
        st.apply_swap(op["q"][0], op["q"][1], op["ctrl"])


def _worker(rank, world, port, task, opts, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import cunqa_b200 as cb
    from oracle import oracle_np as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    plan = cb.plan(task, {**opts, "world": world, "rank": rank})
    nl = plan["n_local"]
    st = O.NumpyState(nl)
    st.v[0] = 0.0
    basis = plan.get("basis", 0)
    if (basis >> nl) == rank:
        st.v[basis & ((1 << nl) - 1)] = 1.0
    n_swaps = 0
    exchanges = []   # the pair lists this rank executes, in order (must be identical on every rank)

    def exchange(pairs):
        # one m-qubit exchange exactly as dist.cu's flush_dist_t issues it: amplitude (rank bits x, local bits y) trades
        # places with (rank bits y, local bits x).  Replayed with an all-gather (the states are tiny here).
        exchanges.append(list(pairs))
        mine = torch.from_numpy(st.v.view(np.float64).copy())
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        allv = [a.numpy().view(np.complex128) for a in allv]
        idx = np.arange(1 << nl)
        src_rank = np.full(1 << nl, rank)
        src_idx = idx.copy()
        for gb, lb in pairs:
            y = (idx >> lb) & 1                      # my amplitude's local bit -> the source's rank bit
            x = (rank >> gb) & 1                     # my rank bit -> the source's local bit
            src_rank = (src_rank & ~(1 << gb)) | (y << gb)
            src_idx = (src_idx & ~(1 << lb)) | (x << lb)
        st.v[:] = np.stack(allv)[src_rank, src_idx]

    group, group_batch = [], None
    max_group = int(os.environ.get("CB200_SWAP_GROUP", "4"))

    def flush_group():
        nonlocal group
        if group:
            exchange(group)
        group = []

    for seg in plan["segments"]:
        if "swap" in seg:
            gbit, lbit = seg["swap"]
            n_swaps += 1
            if seg["batch"] != group_batch:
                flush_group()
            group_batch = seg["batch"]
            group.append((gbit, lbit))
            if len(group) == max_group:
                flush_group()
        else:
            flush_group()
            # execute in the order of the tile passes (the order the GPU uses)
            for ps in seg["passes"]:
                for oi in ps["ops"]:
                    _apply_local(st, seg["ops"][oi])
    flush_group()
    every = [None] * world
    dist.all_gather_object(every, exchanges)
    assert all(e == every[0] for e in every), f"ranks disagree on the exchange sequence: {every}"
    shard = torch.from_numpy(st.v.view(np.float64).copy())
    gathered = [torch.empty_like(shard) for _ in range(world)] if rank == 0 else None
    dist.gather(shard, gathered, dst=0)
    if rank == 0:
        full = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        np.save(out_path, np.concatenate([full.view(np.float64), np.array(plan["l2p"], dtype=np.float64), [float(n_swaps)]]))
    dist.barrier()
    dist.destroy_process_group()


def run_two_ranks(task, n, opts, tmp_path, world=2):
    import torch.multiprocessing as mp
    out = str(tmp_path / "state.npy")
    mp.spawn(_worker, args=(world, _free_port(), task, opts, out), nprocs=world, join=True)
    raw = np.load(out)
    phys = raw[: 2 << n].view(np.complex128)
    l2p = raw[2 << n: (2 << n) + n].astype(int)
    n_swaps = int(raw[-1])
    idx = np.arange(1 << n)
    pidx = np.zeros_like(idx)
    for q in range(n):
        pidx |= ((idx >> q) & 1) << l2p[q]
    return phys[pidx], n_swaps


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_qft_matches_closed_form(lib_built, tmp_path, world):
    from cunqa_b200 import circuits as C
    from oracle import oracle_np as O
    n, x = 9, 0b101100111
    got, n_swaps = run_two_ranks(C.task(C.qft(n, x), n), n, {"tile_bits": 5, "min_run_bits": 2}, tmp_path, world)
    assert np.abs(got - O.qft_amplitudes(n, x, np.arange(1 << n))).max() < 1e-12
    assert n_swaps >= 1  # H on a global qubit cannot be done without an exchange


def test_sharded_random_circuit_matches_oracle(lib_built, tmp_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import random_circuit
    from cunqa_b200 import circuits as C
    from oracle import oracle_np as O
    n = 8
    ins = random_circuit(n, 120, 11)
    ref, _ = O.run_static(ins, n, backend="numpy")
    got, _ = run_two_ranks(C.task(ins, n), n, {"tile_bits": 4, "min_run_bits": 1}, tmp_path, 2)
    assert np.abs(got - ref.amplitudes()).max() < 1e-12


def _exchange_sequence(plan, max_group=4):
    """the exchanges dist.cu's flush_dist_t derives from a per-rank plan (grouping by batch id)"""
    out, group, batch = [], [], None
    for seg in plan["segments"]:
        if "swap" in seg:
            if seg["batch"] != batch and group:
                out.append(group); group = []
            batch = seg["batch"]
            group.append(tuple(seg["swap"]))
            if len(group) == max_group:
                out.append(group); group = []
        elif group:
            out.append(group); group = []
    if group:
        out.append(group)
    return out


def test_exchange_sequence_is_rank_independent_with_global_controls(lib_built):
    """an op controlled by a global qubit is dropped on the ranks whose bit is 0, so the LOCAL segment
    between two swap batches exists on some ranks only; the exchanges must still be the same everywhere."""
    import cunqa_b200 as cb
    from cunqa_b200 import circuits as C
    n = 16
    ins = [{"name": "cx", "qubits": [15, 14]}] + [{"name": "cx", "qubits": [15, q]} for q in range(5, 13)] + [{"name": "h", "qubits": [13]}]
    seqs = []
    for rank in range(8):
        plan = cb.plan(C.task(ins, n), {"world": 8, "rank": rank, "tile_bits": 6, "min_run_bits": 2})
        seqs.append(_exchange_sequence(plan))
    assert all(s == seqs[0] for s in seqs), seqs
    for ex in seqs[0]:   # no repeated global or local bit inside one exchange
        assert len({g for g, _ in ex}) == len(ex) and len({l for _, l in ex}) == len(ex)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sharded_random_circuit_with_global_controls(lib_built, tmp_path, seed):
    """random circuits whose cx / ccx controls (and targets) sit on global qubits, replayed with the GPU's grouping rule"""
    from cunqa_b200 import circuits as C
    from oracle import oracle_np as O
    n, world = 9, 4
    rng = np.random.default_rng(seed)
    ins = [{"name": "h", "qubits": [q]} for q in range(n)]
    for _ in range(60):
        r = rng.integers(0, 5)
        qs = [int(q) for q in rng.permutation(n)[:3]]
        hi = int(rng.integers(n - 2, n))  # a global qubit
        if r == 0:
            ins.append({"name": "cx", "qubits": [hi, qs[0] if qs[0] != hi else qs[1]]})
        elif r == 1:
            a, b = [q for q in qs if q != hi][:2]
            ins.append({"name": "ccx", "qubits": [hi, a, b]})
        elif r == 2:
            ins.append({"name": "h", "qubits": [hi]})
        elif r == 3:
            ins.append({"name": "ry", "qubits": [qs[0]], "params": [float(rng.uniform(0, 6))]})
        else:
            a = qs[0] if qs[0] != hi else qs[1]
            ins.append({"name": "cx", "qubits": [a, hi]})
    ref, _ = O.run_static(ins, n, backend="numpy")
    got, _ = run_two_ranks(C.task(ins, n), n, {"tile_bits": 4, "min_run_bits": 1}, tmp_path, world)
    assert np.abs(got - ref.amplitudes()).max() < 1e-12


def test_diagonal_and_controls_on_global_qubits_need_no_exchange(lib_built):
    import cunqa_b200 as cb
    from cunqa_b200 import circuits as C
    n = 6
    ins = [{"name": "h", "qubits": [0]}, {"name": "cp", "qubits": [0, 5], "params": [0.3]}, {"name": "rz", "qubits": [5], "params": [0.9]},
           {"name": "cx", "qubits": [5, 1]}, {"name": "cz", "qubits": [4, 5]}, {"name": "mcp", "qubits": [5, 4, 2], "params": [0.2]}]
    for rank in range(4):
        plan = cb.plan(C.task(ins, n), {"world": 4, "rank": rank, "tile_bits": 4, "min_run_bits": 1})
        assert all("swap" not in s for s in plan["segments"])
        # the cx controlled by global qubit 5 exists only on the ranks whose bit is 1
        kinds = [op["kind"] for s in plan["segments"] for op in s["ops"]]
        assert ("permx" in kinds) == bool(rank & 2)


# ---- index arithmetic of the multi-qubit exchange kernel (dist.cu: qubit_swap_kernel), restated in numpy ----------
def _insert0(g, pos):
    return ((g >> pos) << (pos + 1)) | (g & ((1 << pos) - 1))


def _exchange_like_the_kernel(shards, n_local, pairs):
    """every rank runs the kernel's loop over w; returns the shards after the in-place exchange"""
    world = len(shards)
    m = len(pairs)
    order = sorted(range(m), key=lambda j: pairs[j][1])
    lbit = [pairs[j][1] for j in order]
    perm = order
    out = [s.copy() for s in shards]
    rest_bits = n_local - m - 1
    half = 1 << rest_bits
    for rank in range(world):
        x = 0
        for j in range(m):
            x |= ((rank >> pairs[j][0]) & 1) << j
        for w in range(((1 << m) - 1) << rest_bits):
            y = x ^ ((w >> rest_bits) + 1)
            r = (w & (half - 1)) | (0 if x < y else half)
            mi = r
            for j in range(m):
                mi = _insert0(mi, lbit[j])
            pi = mi
            for j in range(m):
                mi |= ((y >> perm[j]) & 1) << lbit[j]
                pi |= ((x >> perm[j]) & 1) << lbit[j]
            peer = rank
            for j in range(m):
                peer = (peer & ~(1 << pairs[j][0])) | (((y >> j) & 1) << pairs[j][0])
            out[rank][mi], out[peer][pi] = out[peer][pi], out[rank][mi]
    return out


@pytest.mark.parametrize("world,pairs", [(2, [(0, 3)]), (4, [(1, 0)]), (4, [(0, 4), (1, 2)]), (8, [(2, 1), (0, 5), (1, 3)]), (8, [(1, 4), (2, 0)])])
def test_multi_qubit_exchange_indexing(world, pairs):
    """the kernel's one-step transposition equals the sequence of single (global bit, local bit) swaps, and every
    amplitude is touched at most once (each unordered rank pair splits the index space in halves)"""
    n_local = 6
    g = world.bit_length() - 1
    n = n_local + g
    full = np.arange(1 << n, dtype=np.int64)            # amplitude = its own global index
    want = full.copy()
    for gb, lb in pairs:                                 # sequential reference: swap index bits (n_local + gb) <-> lb
        idx = np.arange(1 << n)
        hi, lo = (idx >> (n_local + gb)) & 1, (idx >> lb) & 1
        swapped = idx ^ (((hi ^ lo) << (n_local + gb)) | ((hi ^ lo) << lb))
        want = want[swapped]
    shards = [full[r << n_local:(r + 1) << n_local] for r in range(world)]
    got = np.concatenate(_exchange_like_the_kernel(shards, n_local, pairs))
    assert np.array_equal(got, want)
