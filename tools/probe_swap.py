"""exchange-kernel probe, single process: one global<->local qubit swap per H on the top qubit(s); GB/s per direction per GPU"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunqa_b200 as cb

def run(n_local, devices):
    g = len(devices).bit_length() - 1
    n = n_local + g
    sv = cb.StateVector(n, "double", devices=devices)
    sv.apply_gate("h", [0]); sv.sync()
    for rep in range(3):
        d0 = sv.dist_stats()
        t0 = time.perf_counter()
        for q in range(n - g, n):
            sv.apply_gate("h", [q])          # non-diagonal on global qubits -> one exchange of g qubits
        sv.sync()
        wall = (time.perf_counter() - t0) * 1e3
        d1 = sv.dist_stats()
        sent = d1["swap_bytes_sent"] - d0["swap_bytes_sent"]
        ms = d1["swap_ms"] - d0["swap_ms"]
        print(json.dumps({"n_local": n_local, "devices": devices, "swaps": d1["swaps"] - d0["swaps"], "GB_sent_per_gpu": sent / 1e9, "swap_ms": ms,
                          "GBps_per_direction_per_gpu": sent / 1e6 / ms if ms else None, "wall_ms": wall}), flush=True)
    sv.close()

if __name__ == "__main__":
    ng = cb.device_count()
    for nl in (28, 30, 32):
        run(nl, [0, 1])
        if ng >= 4:
            run(nl, [0, 1, 2, 3])
        if ng >= 8:
            run(nl, list(range(8)))
