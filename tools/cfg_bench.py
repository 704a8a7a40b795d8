"""Staged-API throughput of BASELINE.json configs 3 and 4 (and LNW/LSW variants) on one GPU, sample-sized.

    python tools/cfg_bench.py 3 [n_pairs] [strings=1] [reps]     AffineNeedlemanWunsch 1000x1000, full traceback
    python tools/cfg_bench.py 4 [n_pairs] [strings=1] [reps]     BandedSmithWaterman band 64, 10k x 10k
    python tools/cfg_bench.py lnw|lsw R Q n_pairs strings reps

Prints one JSON line: library CUDA-event times (fill / backtrack / total), GCUPS over them, and a parity check of the
first pairs against the CPU oracle (test infrastructure; outside every timed region)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from dpx_gpu_genomics_project_b200 import api, synth  # noqa: E402


def main():
    cfg = sys.argv[1]
    if cfg == "3":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
        strings = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
        R = Q = 1000
        img = synth.mutated_fixed_file_bytes(n, R, Q, 0x5EED0003, 0.02, 0.005, 0.005)
        algo, w = api.ANW, dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1)
    elif cfg == "4":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
        strings = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
        R = Q = 10000
        img = synth.mutated_fixed_file_bytes(n, R, Q, 0x5EED0004, 0.05, 0.01, 0.01)
        algo, w = api.BSW, dict(match=3, mismatch=-1, gap_open=-2, band=64)
    else:
        R, Q, n = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        strings = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
        img = synth.mutated_fixed_file_bytes(n, R, Q, 0x5EED0001, 0.05, 0.02, 0.02)
        algo, w = {"lnw": api.LNW, "lsw": api.LSW, "anw": api.ANW}[cfg], dict(match=3, mismatch=-1, gap_open=-2)
        if algo == api.ANW:
            w = dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1)
    import oracle_lib as ol
    blob, pairs = ol.parse_image(img)
    flags = api.OUT_SCORE | api.OUT_END_COORDS | (api.OUT_STRINGS if strings else 0)
    p = api.make_params(algo, flags=flags, **w)
    eng = api.Engine(0)
    b = eng.upload(blob, pairs)
    best = None
    for _ in range(reps):
        b.run(p); b.sync()
        st = b.stats()
        if best is None or st["total_ms"] < best["total_ms"]:
            best = st
    res = b.fetch()
    n_chk = min(n, 16 if R >= 5000 else 64)
    s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs[:n_chk], strings=bool(strings), threads=8)
    ok = bool((res.scores[:n_chk] == s).all())
    if strings:
        ok = ok and res.strings[:n_chk] == t
    if algo in (api.LSW, api.BSW):
        ok = ok and bool((res.end_row_col[:n_chk] == e).all())
    cells = best["cells"]
    out = {"cfg": cfg, "n_pairs": n, "R": R, "Q": Q, "strings": strings, "cells": cells, **{k: best[k] for k in ("fill_ms", "backtrack_ms", "total_ms", "kernel_launches", "kernel_id", "traceback_bytes")},
           "fill_gcups": cells / (best["fill_ms"] * 1e-3) / 1e9 if best["fill_ms"] else None,
           "total_gcups": cells / (best["total_ms"] * 1e-3) / 1e9 if best["total_ms"] else None,
           "tb_gbs": best["traceback_bytes"] / (best["fill_ms"] * 1e-3) / 1e9 if best["fill_ms"] else None,
           "parity_first_pairs": ok}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
