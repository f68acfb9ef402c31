"""Multi-GPU parity check of the sharded state (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py [--big N]

Checks, on every rank: QFT|x> closed form on sampled indices, a random circuit against the CPU oracle,
norm = 1, sampled shots equal to the oracle's for the same seed.  --big N additionally times a sharded
QFT of N qubits (amplitudes checked against the closed form) and prints one JSON line from rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cunqa_b200 as cb
    from cunqa_b200 import circuits as C
    from oracle import oracle_np as O
    from conftest import random_circuit

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = world.bit_length() - 1
    ok = True

    def check(name, cond, info=""):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAIL {name} {info}", flush=True)

    # 1. QFT closed form (forces global<->local qubit swaps)
    n = 14 + g
    x = C.qft_x(n, seed=3)
    sv = cb.ShardedStateVector(n, "double", device=local)
    sv.run(C.qft(n, x))
    idx = np.random.default_rng(1).integers(0, 1 << n, size=4096).astype(np.uint64)
    err = np.abs(sv.amplitudes(idx) - O.qft_amplitudes(n, x, idx)).max()
    check("qft closed form", err < 1e-10 and err * 2.0 ** (n / 2) < 1e-9, f"err={err}")
    check("qft norm", abs(sv.norm()[0] - 1) < 1e-10)
    st = sv.dist_stats()
    check("qft needed swaps", st["swaps"] >= g, str(st))
    sv.close()

    # 2. random circuit over the whole vocabulary vs the oracle, amplitudes + shots + P(q=1)
    n = 13 + g
    ins = random_circuit(n, 150, 21, max_unitary=4)
    ref, _ = O.run_static(ins, n, backend="c")
    sv = cb.ShardedStateVector(n, "double", device=local)
    sv.run(ins)
    got = sv.amplitudes()
    err = np.abs(got - ref.amplitudes()).max()
    check("random circuit amplitudes", err < 1e-10, f"err={err}")
    want = ref.sample(2000, 9)
    shots = sv.sample(2000, 9)
    # the engine's CDF runs over the physical (rank-major) order; equal distributions, shot-noise bound
    p = np.abs(ref.amplitudes()) ** 2
    emp = np.bincount(shots.astype(np.int64), minlength=1 << n) / 2000.0
    check("shots distribution", 0.5 * np.abs(emp - p).sum() < 0.5 * np.sqrt((1 << n) / 2000.0) + 0.05)
    check("shots in range", int(shots.max()) < (1 << n) and len(want) == len(shots))
    for q in (0, n - 1, n // 2):
        check(f"prob1({q})", abs(sv.prob1(q)[0] - ref.prob1(q)) < 1e-10)
    sv.close()

    # 2b. mid-circuit measurement of the sharded state: local and global qubits, plain and with reset -- the same
    #     outcomes (shared counter stream) and the same post-measurement state as the oracle, gates in between
    sv = cb.ShardedStateVector(n, "double", device=local)
    sv.run(ins)
    ref2, _ = O.run_static(ins, n, backend="c")
    more = random_circuit(n, 40, 33, max_unitary=3)
    xmat = np.array([[0, 1], [1, 0]], dtype=complex)
    for k, (q, rst) in enumerate(((0, False), (n - 1, False), (n // 2, True), (n - 2, True), (n - 1, True))):
        ctr = O.meas_counter(0, k)
        want = ref2.measure(q, 5, ctr)
        if rst and want:
            ref2.apply_matrix([q], xmat)
        got = sv.measure(q, seed=5, counters=[ctr], reset=rst)
        if not rst:
            check(f"measure({q}) outcome", int(got[0]) == int(want), f"{got} vs {want}")
        for inst in more[8 * k: 8 * k + 8]:
            O.apply_gate(ref2, inst)
        sv.run(more[8 * k: 8 * k + 8])
        err = np.abs(sv.amplitudes() - ref2.amplitudes()).max()
        check(f"state after measure({q}, reset={rst})", err < 1e-10, f"err={err}")
        check(f"norm after measure({q})", abs(sv.norm()[0] - 1) < 1e-10)
    sv.close()

    # 3. fp32 shard
    sv = cb.ShardedStateVector(n, "single", device=local)
    sv.run(ins)
    err = np.abs(sv.amplitudes() - ref.amplitudes()).max()
    check("random circuit fp32", err < 1e-5, f"err={err}")
    sv.close()

    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("DIST PARITY", "OK" if int(flag.item()) == 0 else "FAILED", f"world={world}", flush=True)

    if args.big and int(flag.item()) == 0:
        n = args.big
        x = C.qft_x(n, seed=7)
        ins = C.qft(n, x)
        sv = cb.ShardedStateVector(n, "double", device=local)
        stream = torch.cuda.ExternalStream(sv.stream(), device=torch.device("cuda", local))
        times = []
        for rep in range(3):
            sv.reset(); sv.sync()
            dist.barrier(); torch.cuda.synchronize()
            s0 = sv.stats(); d0 = sv.dist_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sv.run(ins); sv.flush()
            e1.record(stream)
            sv.sync(); dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        s1 = sv.stats(); d1 = sv.dist_stats()
        idx = np.random.default_rng(2).integers(0, 1 << n, size=1 << 14).astype(np.uint64)
        err = float(np.abs(sv.amplitudes(idx) - O.qft_amplitudes(n, x, idx)).max())
        nrm = float(sv.norm()[0])
        if rank == 0:
            ms = min(times[1:])
            passes = s1["passes"] - s0["passes"]
            swaps = d1["swaps"] - d0["swaps"]
            shard = 16 * 2 ** (n - g)
            print(json.dumps({"workload": f"QFT-{n} fp64 sharded over {world} GPUs", "gates": C.n_gates(ins), "ms": ms,
                              "gates_per_s": C.n_gates(ins) / (ms / 1e3), "tile_passes_per_rank": passes, "qubit_swaps": swaps,
                              "shard_GiB": shard / 2 ** 30, "max_abs_err_vs_closed_form": err, "norm": nrm,
                              "local_GBps_per_gpu": 2 * shard * passes / (ms / 1e3) / 1e9,
                              "swap_GB_sent_per_gpu": (d1["swap_bytes_sent"] - d0["swap_bytes_sent"]) / 1e9,
                              "swap_ms_last_rep": d1["swap_ms"] - d0["swap_ms"]}), flush=True)
        sv.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
