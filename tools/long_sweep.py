"""Separates per-row cost from pipeline-fill cost of the long-pair kernel: time(Q) at fixed R, per lane width K."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dpx_gpu_genomics_project_b200 import api, synth, longpair
R = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
eng = api.Engine(0)
p = api.make_params(api.LSW)
rng = synth.Rng(5)
ref = synth.random_seq(rng, R)
for K in sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"]:
    eng.set_option("long_k", int(K))
    row = []
    for Q in (25_000, 50_000, 100_000, 200_000, 400_000):
        qry = synth.random_seq(rng, Q)
        job = longpair.StripedLongPair(eng, p, ref, qry, 0, 1, None)
        job.run(); res, ms = job.run()
        job.free()
        row.append((Q, round(ms, 2)))
    (q1, t1), (q2, t2) = row[-2], row[-1]
    per_row_us = (t2 - t1) * 1e3 / (q2 - q1)
    fill_ms = t2 - per_row_us * q2 / 1e3
    nw = (R + 32 * int(K) - 1) // (32 * int(K))
    print(json.dumps({"R": R, "K": int(K), "warps": nw, "times_ms": row, "per_row_us": round(per_row_us, 4), "fill_ms": round(fill_ms, 2),
                      "fill_us_per_warp": round(fill_ms * 1e3 / nw, 2), "steady_gcups": round(R / per_row_us / 1e3, 1)}))
