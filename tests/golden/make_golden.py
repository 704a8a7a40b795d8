"""Regenerates the golden fixtures: inputs from the seeded generator, expected text from the
COMPILED, UNMODIFIED reference classes (oracle/_ref/ref_align, built by oracle/Makefile from
/root/reference/c++ in place).  Run in the build container (needs /root/reference):

    make -C oracle && python tests/golden/make_golden.py

Outputs (committed): tests/golden/<name>.in.txt, tests/golden/<name>.<ALGO>.out.txt, tests/golden/ties.pairs.txt + ties.LSW_ALL.out.txt (the
reference built with -DBACKTRACK_ALL, oracle/_ref/ref_align_all),
tests/golden/bsw_python_scores.json (scores of the reference's Python banded prototype).
Scoring = the reference's golden parameters (correct-outputs/LNW/web-scraper-LNW.py:139-141,
correct-outputs/ANW/web-scraper-ANW.py): match 3, mismatch -1, gap -2; Gotoh open -3, extend -1.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle_lib as ol  # noqa: E402
from dpx_gpu_genomics_project_b200 import synth  # noqa: E402


def adversarial(seed):
    """Tie-heavy / degenerate pairs: lengths 0..40, alphabets {0},{0,1},{0..3},{0..4}; empty ref/query/both."""
    rng = synth.Rng(seed)
    pairs = [(b"", b""), (b"0", b""), (b"", b"1"), (b"0", b"0"), (b"0", b"1"), (b"00000", b"00000"),
             (b"0101010101", b"1010101010"), (b"0123", b"3210"), (b"4", b"4"), (b"0404", b"4040")]
    for alpha in (b"0", b"01", b"0123", b"01234"):
        for _ in range(60):
            R = int(rng.below(1, 41)[0]); Q = int(rng.below(1, 41)[0])
            r = synth.random_seq(rng, R, alpha)
            if rng.uniform(1)[0] < 0.5:
                q = synth.mutate(rng, r, 0.1, 0.05, 0.05, alpha)[:40]
            else:
                q = synth.random_seq(rng, Q, alpha)
            pairs.append((r, q))
    return pairs


def cfg1_like(seed, n, lo, hi, dmax):
    """SURVEY §8d config-1 substitute: R~U[lo,hi], query = ref mutated 5% sub, 2% ins, 2% del."""
    rng = synth.Rng(seed)
    pairs = []
    for _ in range(n):
        R = lo + int(rng.below(1, hi - lo + 1)[0])
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05, 0.02, 0.02)
        pairs.append((r, q))
    return pairs


def shapes(seed):
    """Q>>R, R>>Q, lengths straddling 32/64/128 boundaries, independent random pairs."""
    rng = synth.Rng(seed)
    pairs = []
    for (R, Q) in [(31, 31), (32, 32), (33, 33), (63, 65), (64, 64), (65, 63), (127, 129), (128, 128), (129, 127),
                   (5, 300), (300, 5), (1, 200), (200, 1), (257, 255), (150, 150), (150, 150), (513, 40), (40, 513)]:
        pairs.append((synth.random_seq(rng, R), synth.random_seq(rng, Q)))
        r = synth.random_seq(rng, R)
        pairs.append((r, (synth.mutate(rng, r, 0.05, 0.02, 0.02) + synth.random_seq(rng, Q))[:Q]))
    return pairs


def ties(seed):
    """Pairs with SEVERAL maximum cells (periodic sequences, repeated cores, tiny alphabets), all scoring > 0: the input of the
    BACKTRACK_ALL golden (score-0 pairs are left out: there the reference's all-maxima mode walks uninitialised cells)."""
    rng = synth.Rng(seed)
    pairs = [(b"0101", b"1010"), (b"0123012", b"0123"), (b"00100", b"00"), (b"0101010101", b"01010"), (b"012012012", b"12")]
    for alpha in (b"0", b"01", b"012", b"0123"):
        for k in range(40):
            R = 1 + int(rng.below(1, 40)[0]); Q = 1 + int(rng.below(1, 40)[0])
            if k % 3 == 0:
                unit = synth.random_seq(rng, 1 + int(rng.below(1, 4)[0]), alpha)
                r, q = (unit * 40)[:R], (unit * 40)[1:1 + Q]
            elif k % 3 == 1:
                r, q = synth.random_seq(rng, R, alpha), synth.random_seq(rng, Q, alpha)
            else:
                core = synth.random_seq(rng, 2 + int(rng.below(1, 6)[0]), alpha)
                r = synth.random_seq(rng, R // 3, alpha) + core + synth.random_seq(rng, R // 3, alpha) + core
                q = core + synth.random_seq(rng, Q // 2, alpha) + core
            if set(r) & set(q):
                pairs.append((r, q))
    return pairs


SETS = {
    "adversarial": lambda: adversarial(0x5EED0000 + 101),
    "cfg1_small": lambda: cfg1_like(0x5EED0000 + 1, 120, 100, 300, 20),
    "shapes": lambda: shapes(0x5EED0000 + 102),
    "mid": lambda: cfg1_like(0x5EED0000 + 103, 6, 900, 1300, 0),
}


def python_banded_scores():
    """Scores of python/LinearBandedSmithWaterman.py (the only runnable banded implementation in the
    reference) with BAND = W+1 (its range :71 is |i-j| <= BAND-1).  Only initializeMemoMatrix +
    performRecursiveAnalysis are called (execute() enumerates all optimal paths)."""
    sys.path.insert(0, "/root/reference/python")
    import io
    import contextlib
    from LinearBandedSmithWaterman import LinearBandedSmithWatermanAligner  # type: ignore
    rng = synth.Rng(0x5EED0000 + 104)
    cases = []
    for k in range(60):
        R = 5 + int(rng.below(1, 90)[0])
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.08, 0.04, 0.04) if k % 3 else synth.random_seq(rng, 5 + int(rng.below(1, 90)[0]))
        W = [0, 1, 2, 3, 5, 8, 16, 64][k % 8]
        with contextlib.redirect_stdout(io.StringIO()):
            a = LinearBandedSmithWatermanAligner(r.decode(), q.decode(), 3, -1, -2, W + 1)
            a.initializeMemoMatrix()
            a.performRecursiveAnalysis()
        cases.append({"ref": r.decode(), "qry": q.decode(), "band": W, "score": int(np.max(a.Memo))})
    return cases


def main():
    assert ol.have_ref_binary(), "build oracle/_ref first: make -C oracle"
    for name, fn in SETS.items():
        img = synth.pairs_to_file_bytes(fn())
        path = os.path.join(HERE, f"{name}.in.txt")
        with open(path, "wb") as f:
            f.write(img)
        for algo, kw in ((ol.LNW, dict(gap_open=-2)), (ol.LSW, dict(gap_open=-2)), (ol.ANW, dict(gap_open=-3, gap_extend=-1))):
            out = ol.run_reference(algo, path, **kw)
            with open(os.path.join(HERE, f"{name}.{ol.ALGO_NAMES[algo]}.out.txt"), "wb") as f:
                f.write(out)
            print(name, ol.ALGO_NAMES[algo], len(out), "bytes")
    # LinearSmithWaterman built with -DBACKTRACK_ALL (c++/LinearSmithWaterman.h:9): oracle/_ref/ref_align_all
    img = synth.pairs_to_file_bytes(ties(0x5EED0000 + 105))
    path = os.path.join(HERE, "ties.pairs.txt")
    with open(path, "wb") as f:
        f.write(img)
    out = ol.run_reference_all(path, 3, -1, -2)
    with open(os.path.join(HERE, "ties.LSW_ALL.out.txt"), "wb") as f:
        f.write(out)
    print("ties LSW_ALL", len(out), "bytes")
    with open(os.path.join(HERE, "bsw_python_scores.json"), "w") as f:
        json.dump(python_banded_scores(), f, indent=0)
    print("done")




def fakedpx_vectors():
    """Known-answer vectors of the reference's only unit test (c++/testFakeDPX.cpp:10-113), extracted as
    data: [{"fn": "__vimax3_s32", "args": [1,2,3], "want": 3, "preds": {"pred_hi": false, ...}}, ...]."""
    import re
    src = open("/root/reference/c++/testFakeDPX.cpp").read()
    out = []
    for m in re.finditer(r"assert\(FakeDPX::(\w+)\(([^)]*)\)\s*==\s*([-0-9a-fA-FxX]+)((?:\s*&&\s*!?\w+)*)\);", src):
        fn, args, want, preds = m.groups()
        vals = [int(a.strip(), 0) for a in args.split(",") if not a.strip().startswith("&")]
        pm = {}
        for t in re.findall(r"&&\s*(!?)(\w+)", preds):
            pm[t[1]] = (t[0] != "!")
        out.append({"fn": fn, "args": vals, "want": int(want, 0), "preds": pm})
    return out


if __name__ == "__main__":
    main()
    with open(os.path.join(HERE, "fakedpx_vectors.json"), "w") as f:
        json.dump(fakedpx_vectors(), f, indent=0)
    print("fakedpx vectors:", len(fakedpx_vectors()))
