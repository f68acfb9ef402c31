"""config 5 host/device time split: CB200_TRACE=1 python tools/probe_config5.py [groups]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunqa_b200 as cb
from cunqa_b200 import circuits as C
G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
be = cb.Backend(0)
texts = [json.dumps(m) for m in C.config5_groups(G, n_each=11, layers=2, shots=1024, seed=2024)]
for rep in range(3):
    t0 = time.perf_counter(); be.execute_many(texts, parse=False); print("qc groups round", round((time.perf_counter() - t0) * 1e3, 2), "ms", be.batch_info(), flush=True)
n, layers = 24, 2
be.execute_many([json.dumps(C.task(C.hea_ansatz(n, layers, seed=7 + g), n, shots=1024, seed=g)) for g in range(G)], parse=False)
rng = np.random.default_rng(0)
upd = [json.dumps({"params": list(rng.uniform(0, 6.28, 2 * n * layers)), "shots": 1024}) for _ in range(G)]
for rep in range(3):
    t0 = time.perf_counter(); be.execute_many(upd, parse=False); print("vqe update round", round((time.perf_counter() - t0) * 1e3, 2), "ms", be.batch_info(), flush=True)
