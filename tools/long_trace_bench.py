"""BASELINE config 5 with its alignment: 1 Mbp x 1 Mbp LinearSmithWaterman, score + end cell + the three printed lines, on one GPU.
Prints one JSON line with the stage times (device ms) and the structural checks of the result (no full-matrix oracle exists at
this size: the score / end cell are checked against the rolling-row oracle only when --oracle is given, ~20 min of CPU).
usage: python tools/long_trace_bench.py [R] [Q] [--oracle] [--mut=sub,ins,del]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from dpx_gpu_genomics_project_b200 import api, synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    R = int(args[0]) if args else 1_000_000
    Q = int(args[1]) if len(args) > 1 else R
    rng = synth.Rng(0x5EED0005)
    ref = synth.random_seq(rng, R)
    mut = (0.01, 0.001, 0.001)
    for a in sys.argv[1:]:
        if a.startswith("--mut="):
            mut = tuple(float(x) for x in a[6:].split(","))
    qry = synth.mutate(rng, ref, *mut)
    qry = (qry + synth.random_seq(rng, Q))[:Q]
    eng = api.Engine(0)
    p = api.make_params(api.LSW)
    out = {"R": R, "Q": Q, "mutation": list(mut)}
    for rep in range(2):                       # first call pays the 12 GB of cudaMalloc page mapping
        t0 = time.perf_counter()
        end, start, lines, st = eng.align_long_pair_strings(p, ref, qry)
        wall = time.perf_counter() - t0
        out[f"run{rep}"] = dict(wall_s=round(wall, 4), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
    t0 = time.perf_counter(); plain = eng.align_long_pair(p, ref, qry); out["score_only_wall_s"] = round(time.perf_counter() - t0, 4)
    rel = np.frombuffer(lines[1], dtype=np.uint8)
    rescored = 3 * int((rel == ord("*")).sum()) - int((rel == ord("|")).sum()) - 2 * int((rel == ord(" ")).sum())
    out.update(result=list(end), start=list(start), line_len=len(lines[0]), rescored=rescored,
               checks=dict(same_as_score_only=tuple(end) == tuple(plain), rescored_equals_score=rescored == end[0],
                           ref_line_spells_ref=lines[0].replace(b"_", b"") == ref[start[1]:end[2]],
                           qry_line_spells_qry=lines[2].replace(b"_", b"") == qry[start[0]:end[1]]))
    st = out["run1"]
    out["gcups_with_alignment"] = round(R * Q / (st["fwd_ms"] + st["walk_ms"]) / 1e6, 1)
    if "--oracle" in sys.argv:
        import oracle_lib as ol
        out["oracle"] = list(ol.lsw_score_only(ol.params(ol.LSW), ref, qry))
    print(json.dumps(out))
    assert all(out["checks"].values()), out["checks"]


if __name__ == "__main__":
    main()
