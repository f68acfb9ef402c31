"""CPU: the host fuser + tile-pass scheduler (H1) must be an exact re-arrangement of the circuit.

The plan is obtained through the C ABI (cb200_plan_json, no GPU needed) and replayed with the
numpy oracle; the result must equal the gate-by-gate oracle run of the original circuit.
"""
import numpy as np
import pytest

import cunqa_b200 as cb
from cunqa_b200 import circuits as C
from oracle import oracle_np as O
from conftest import random_circuit

X = np.array([[0, 1], [1, 0]], dtype=complex)


def replay_plan(plan, n):
    st = O.make_state(n, "numpy" if n <= 12 else "c")
    if plan.get("basis", 0):  # leading X gates on |0...0> are folded into the initial basis state
        st.amplitudes()[0] = 0.0
        st.amplitudes()[plan["basis"]] = 1.0
    executed = []
    for ps in plan["passes"]:
        tile = set(ps["tile_q"])
        assert len(ps["tile_q"]) == len(tile) <= 13
        assert ps["tile_q"] == sorted(ps["tile_q"]) and ps["tile_q"][:ps["L"]] == list(range(ps["L"]))
        for oi in ps["ops"]:
            op = plan["ops"][oi]
            executed.append(oi)
            m = np.array([complex(*z) for z in op["m"]])
            if op["kind"] == "dense":
                assert set(op["q"]) <= tile, "dense targets must be staged in the tile"
                d = 1 << len(op["q"])
                st.apply_matrix(op["q"], m.reshape(d, d), op["ctrl"])
            elif op["kind"] == "diag":
                st.apply_diagonal(op["q"], m)
            elif op["kind"] == "permx":
                assert set(op["q"]) <= tile
                st.apply_matrix(op["q"], X, op["ctrl"])
            else:
                assert set(op["q"]) <= tile
                st.apply_swap(op["q"][0], op["q"][1], op["ctrl"])
    assert sorted(executed) == list(range(len(plan["ops"]))), "every fused op runs exactly once"
    phys = st.amplitudes()
    l2p = plan["l2p"]
    idx = np.arange(1 << n)
    pidx = np.zeros_like(idx)
    for q in range(n):
        pidx |= ((idx >> q) & 1) << l2p[q]
    return phys[pidx]


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("opts", [None, {"tile_bits": 5, "min_run_bits": 2}, {"tile_bits": 7, "min_run_bits": 3, "max_pass_cost": 6},
                                  {"fuse": False, "tile_bits": 6, "min_run_bits": 0}])
def test_plan_equals_circuit(lib_built, seed, opts):
    n = 4 + seed % 6
    ins = random_circuit(n, 90, seed)
    plan = cb.plan(C.task(ins, n), opts)
    ref, _ = O.run_static(ins, n, backend="numpy")
    got = replay_plan(plan, n)
    assert np.abs(got - ref.amplitudes()).max() < 1e-12
    assert plan["gates"] == sum(1 for i in ins if i["name"] not in ("id",))


def test_qft_plan_is_few_passes(lib_built):
    n = 30
    plan = cb.plan(C.task(C.qft(n, 0), n))
    assert plan["relabelled_swaps"] == n // 2       # uncontrolled swaps cost no pass
    px = cb.plan(C.task(C.qft(n, 0b1011 << 20), n))
    assert px["basis"] == 0b1011 << 20 and len(px["ops"]) == len(plan["ops"])  # X-prep costs no op at all
    assert len(plan["passes"]) <= 6                  # vs 480 per-gate sweeps
    # all cp gates were folded into diagonal tables of <= 10 qubits
    assert all(len(o["q"]) <= 10 for o in plan["ops"] if o["kind"] == "diag")


def test_qft_plan_small_exact(lib_built):
    n = 11
    x = 1234
    plan = cb.plan(C.task(C.qft(n, x), n), {"tile_bits": 6, "min_run_bits": 2})
    got = replay_plan(plan, n)
    assert np.abs(got - O.qft_amplitudes(n, x, np.arange(1 << n))).max() < 1e-12


def test_qv_plan(lib_built):
    n = 14
    ins = C.quantum_volume(n, seed=3)
    plan = cb.plan(C.task(ins, n), {"tile_bits": 8, "min_run_bits": 3})
    ref, _ = O.run_static(ins, n, backend="c")
    assert np.abs(replay_plan(plan, n) - ref.amplitudes()).max() < 1e-12
    # the point of tile passes: several 2-qubit blocks per HBM round trip
    assert len(plan["passes"]) < len(ins) / 2


def test_hea_fuses_single_qubit_gates(lib_built):
    n = 10
    ins = C.hea_ansatz(n, layers=3, measure=False)
    plan = cb.plan(C.task(ins, n))
    # ry/rz are absorbed into the neighbouring cx blocks
    assert len(plan["ops"]) <= 3 * (n - 1) + n
    ref, _ = O.run_static(ins, n, backend="numpy")
    assert np.abs(replay_plan(plan, n) - ref.amplitudes()).max() < 1e-12


def test_bad_instruction_reports_error(lib_built):
    from cunqa_b200.binding import B200Error
    with pytest.raises(B200Error, match="unsupported instruction"):
        cb.plan(C.task([{"name": "hz2", "qubits": [0]}], 2))
    with pytest.raises(B200Error, match="out of range"):
        cb.plan(C.task([{"name": "h", "qubits": [5]}], 2))
    with pytest.raises(B200Error, match="duplicate"):
        cb.plan(C.task([{"name": "cx", "qubits": [1, 1]}], 2))


@pytest.mark.parametrize("kind", ["qv", "random", "hea"])
def test_lookahead_tile_choice_is_equivalent_and_needs_fewer_passes(lib_built, kind):
    """the tile qubits of a pass are chosen by look-ahead (by default for states of >= 2^24 amplitudes): same state, fewer passes"""
    n = 20
    ins = {"qv": lambda: C.quantum_volume(n, depth=6, seed=5), "random": lambda: random_circuit(n, 120, 17),
           "hea": lambda: C.hea_ansatz(n, layers=3, measure=False)}[kind]()
    opts = {"tile_bits": 9, "min_run_bits": 3, "lookahead_min_qubits": 20}
    look = cb.plan(C.task(ins, n), opts)
    first_come = cb.plan(C.task(ins, n), dict(opts, lookahead=0))
    assert len(look["ops"]) == len(first_come["ops"])
    assert len(look["passes"]) <= len(first_come["passes"])
    if kind != "random":
        assert len(look["passes"]) < len(first_come["passes"])
    ref, _ = O.run_static(ins, n, backend="c")
    assert np.abs(replay_plan(look, n) - ref.amplitudes()).max() < 1e-12
    assert np.abs(replay_plan(first_come, n) - ref.amplitudes()).max() < 1e-12


def test_lookahead_pass_counts_at_the_headline_sizes(lib_built):
    qv = C.task(C.quantum_volume(33, seed=2024), 33)
    assert len(cb.plan(qv)["passes"]) <= 66 < 92 <= len(cb.plan(qv, {"lookahead": 0})["passes"])
    hea = C.task(C.hea_ansatz(30, 4, 1, measure=False), 30)
    assert len(cb.plan(hea)["passes"]) <= 6


def test_phase_fan_takes_the_commuting_ladder_terms_of_a_qft_pass(lib_built):
    """encode_pass pulls the controlled-phase terms that commute with the rest of their pass into one fan op: a QFT-30
    pass then carries a handful of small tables (the ladder terms among its own six qubits) instead of up to 17"""
    n = 30
    t = C.task(C.qft(n, C.qft_x(n, seed=7)), n)
    plain = cb.plan(t, {"fan": False})["passes"]
    fanned = cb.plan(t, {"fan": True})["passes"]
    assert len(plain) == len(fanned)
    assert max(p["diag_tables"] for p in plain) >= 14
    for p, f in zip(plain, fanned):
        if p["diag_tables"] >= 5:
            assert "fan" in f and f["diag_tables"] <= 4
            assert 1 <= f["fan"]["centers"] <= 8 and f["fan"]["leaf_bits_in_tile"] <= 6 and f["fan"]["leaf_chunks_outside"] <= 4
        else:
            assert "fan" not in f
    # dense-only circuits never get one
    assert all("fan" not in p for p in cb.plan(C.task(C.quantum_volume(26, depth=4, seed=1), 26), {"fan": True})["passes"])


def test_fuser_absorbs_both_rotations_into_the_entangler(lib_built):
    """ry(a), ry(b), cx(a, b): the entangler swallows the rotation next to it and, widened, goes on to the other one
    (round 2: a merge that widens the earlier op keeps looking further back)"""
    ins = [{"name": "ry", "qubits": [0], "params": [0.3]}, {"name": "ry", "qubits": [1], "params": [1.1]},
           {"name": "h", "qubits": [2]}, {"name": "cx", "qubits": [0, 1]}]
    plan = cb.plan(C.task(ins, 3))
    assert len(plan["ops"]) == 2                       # the block on (0, 1) and the h on 2
    ref, _ = O.run_static(ins, 3, backend="numpy")
    assert np.abs(replay_plan(plan, 3) - ref.amplitudes()).max() < 1e-12
    # a gate in between that does not commute stops the search
    ins2 = [{"name": "ry", "qubits": [0], "params": [0.3]}, {"name": "cx", "qubits": [0, 2]}, {"name": "ry", "qubits": [1], "params": [1.1]},
            {"name": "cx", "qubits": [0, 1]}]
    plan2 = cb.plan(C.task(ins2, 3))
    ref2, _ = O.run_static(ins2, 3, backend="numpy")
    assert np.abs(replay_plan(plan2, 3) - ref2.amplitudes()).max() < 1e-12


def test_permutation_runs_and_block_forms_in_the_plan(lib_built):
    """device-op letters of the plan: a run of CX gates is one linear permutation (L<gates>), the register blocks of a lean
    pass are widened (P = two 2-qubit gates on four bits, P(221) / P(200) with spectator bits)"""
    n = 20
    kinds = [p["device_kinds"] for p in cb.plan(C.task(C.ghz(n, measure=False), n))["passes"]]
    assert any("L" in k for k in kinds), kinds
    off = [p["device_kinds"] for p in cb.plan(C.task(C.ghz(n, measure=False), n), {"lin_perm": False})["passes"]]
    assert not any("L" in k for k in off) and sum(k.count("X") for k in off) > sum(k.count("X") for k in kinds)
    qv = cb.plan(C.task(C.quantum_volume(16, seed=3), 16))
    for p in qv["passes"]:
        assert set(p["device_kinds"]) <= set("P(0123456789)"), p["device_kinds"]   # every op of a QV pass is a register block
        assert all(len(bits) == 4 for bits in p["pair_bits"])                       # ... on four local bits


def test_linear_permutation_tables_equal_the_gates_applied_one_by_one(lib_built):
    """OP_LINPERM (round 2): the host composes a run of X / CX / SWAP gates into one GF(2)-affine map of the tile index and
    folds the shared-memory swizzle into its columns.  Bit-exact check on the CPU: for every destination index of the tile
    the table's source equals the index obtained by undoing the gates of the run one by one."""
    import re
    rng = np.random.default_rng(3)
    n = 9                                            # one tile holds the whole register: tile bit = qubit
    ins = [{"name": "h", "qubits": [0]}]
    for _ in range(40):
        a, b = map(int, rng.permutation(n)[:2])
        r = int(rng.integers(0, 4))
        # (uncontrolled swaps are relabelled away by the queue and never reach a pass: cswap would not be linear)
        ins.append({"name": "x", "qubits": [a]} if r == 0 else {"name": "cx", "qubits": [a, b]})
    ins.append({"name": "h", "qubits": [1]})
    plan = cb.plan(C.task(ins, n), {"fuse": False, "tile_bits": n, "min_run_bits": 0})
    assert plan["l2p"] == list(range(n))
    checked = 0
    for p in plan["passes"]:
        assert p["tile_q"] == list(range(n))
        host = [plan["ops"][i] for i in p["ops"]]
        tokens = re.findall(r"L\d+|P\(\d+\)|P|[DXSTF]\d*", p["device_kinds"])
        pos, tables = 0, iter(p["lin_perms"])
        for tok in tokens:
            if tok.startswith("L"):
                k = int(tok[1:])
                run = host[pos:pos + k]
                lin = next(tables)
                swz = lambda i: i ^ ((i >> 3) & 7)   # fp64 tiles; an involution
                for d in range(1 << n):
                    slot = lin[15]
                    for i in range(n):
                        if (d >> i) & 1:
                            slot ^= lin[i]
                    src = d                          # undo the run: last gate first (every gate is its own inverse)
                    for g in reversed(run):
                        assert g["kind"] in ("permx", "swap"), g
                        if g["kind"] == "permx":
                            if all((src >> c) & 1 for c in g["ctrl"]):
                                src ^= 1 << g["q"][0]
                        else:
                            x, y = g["q"]
                            if ((src >> x) & 1) != ((src >> y) & 1):
                                src ^= (1 << x) | (1 << y)
                    assert swz(slot) == src, (d, slot, src)
                pos += k
                checked += 1
            elif tok.startswith("P("):
                pos += 1 if tok[2:5].endswith("00") else 2
            elif tok == "P":
                pos += 2
            else:
                pos += 1
        assert pos == len(host)
    assert checked >= 1


def test_tile_number_to_base_segment_masks(lib_built):
    """PassParams::tile_seg (round 2): the kernel turns a tile number into the tile's base address with
    sum_k (x & seg[k]) << k; it must equal inserting a zero bit at every high tile qubit, lowest first"""
    rng = np.random.default_rng(11)
    for n, ins, opts in ((22, C.quantum_volume(22, depth=3, seed=1), None), (20, C.qft(20, 5), {"tile_bits": 9, "min_run_bits": 3}),
                         (17, C.quantum_volume(17, depth=2, seed=2), {"tile_bits": 6, "min_run_bits": 0})):
        for p in cb.plan(C.task(ins, n), opts)["passes"]:
            tq, L, seg = p["tile_q"], p["L"], p["tile_seg"]
            t = len(tq)
            assert len(seg) == t - L + 1 and tq[:L] == list(range(L))
            for x in [0, (1 << (n - t)) - 1] + [int(v) for v in rng.integers(0, 1 << (n - t), 50)]:
                want = x
                for q in tq[L:]:                      # ascending: insert a zero at bit (q - L)
                    pos = q - L
                    want = ((want >> pos) << (pos + 1)) | (want & ((1 << pos) - 1))
                got = 0
                for k, m in enumerate(seg):
                    got |= (x & m) << k
                assert got == want, (tq, L, x)
                base = got << L
                assert all(not (base >> q) & 1 for q in tq)   # the tile's own qubits are zero in its base


def test_register_block_descriptors(lib_built):
    """OP_PAIR descriptors as the kernel reads them: the group index g of a thread is deposited around the block's local
    bits by sum_m (g & seg[m]) << m -- a bijection onto the tile indices whose local bits are zero -- and sboff[m] is the
    swizzled tile index of local bit m; lean passes hold four local bits (spectators included)"""
    swz = lambda i: i ^ ((i >> 3) & 7)
    seen = 0
    for n, ins in ((16, C.quantum_volume(16, seed=4)), (14, C.hea_ansatz(14, 3, measure=False))):
        for p in cb.plan(C.task(ins, n))["passes"]:
            t = len(p["tile_q"])
            for d in p["pair_desc"]:
                bits, R = d["tbit"], len(d["tbit"])
                assert R == 4 and len(set(bits)) == R and all(0 <= b < t for b in bits)
                assert d["sboff"] == [swz(1 << b) for b in bits]
                mask = sum(1 << b for b in bits)
                idx = set()
                for g in range(1 << (t - R)):
                    i = 0
                    for m, sm in enumerate(d["seg"]):
                        i |= (g & sm) << m
                    assert i & mask == 0 and i < (1 << t)
                    idx.add(i)
                assert len(idx) == 1 << (t - R)
                seen += 1
    assert seen > 10
