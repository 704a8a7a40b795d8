"""Per-row step time of the long-pair kernel as a function of how many warps share an SM (K fixed)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpx_gpu_genomics_project_b200 import api, synth, longpair
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eng = api.Engine(0)
eng.set_option("long_k", K)
p = api.make_params(api.LSW)
rng = synth.Rng(5)
for nw in (1, 4, 148, 592, 1184, 2368):
    R = nw * 32 * K
    ref = synth.random_seq(rng, R)
    ts = []
    for Q in (100_000, 200_000):
        qry = synth.random_seq(rng, Q)
        job = longpair.StripedLongPair(eng, p, ref, qry, 0, 1, None)
        job.run(); res, ms = job.run(); job.free()
        ts.append(ms)
    per_row_us = (ts[1] - ts[0]) * 1e3 / 100_000
    print(json.dumps({"K": K, "warps": nw, "per_row_us": round(per_row_us, 4), "cycles_per_step@1965": round(per_row_us * 1965), "ms": [round(t, 2) for t in ts]}))
