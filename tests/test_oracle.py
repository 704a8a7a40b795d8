"""The CPU oracle (oracle/dpx_oracle.c) against (1) the committed golden fixtures produced by the
compiled, unmodified reference classes, (2) the live reference binary when it is present,
(3) the pins of the repaired banded semantics."""
import glob
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(GOLD, "*.in.txt")))
KW = {ol.LNW: dict(gap_open=-2), ol.LSW: dict(gap_open=-2), ol.ANW: dict(gap_open=-3, gap_extend=-1)}


def oracle_text(algo, img, **kw):
    blob, pairs = ol.parse_image(img)
    scores, _, strs = ol.align_batch(ol.params(algo, **kw), blob, pairs)
    return ol.format_text(scores, strs)


@pytest.mark.parametrize("name", SETS)
@pytest.mark.parametrize("algo", [ol.LNW, ol.LSW, ol.ANW])
def test_oracle_matches_golden(name, algo):
    img = open(os.path.join(GOLD, f"{name}.in.txt"), "rb").read()
    want = open(os.path.join(GOLD, f"{name}.{ol.ALGO_NAMES[algo]}.out.txt"), "rb").read()
    assert oracle_text(algo, img, **KW[algo]) == want


@pytest.mark.skipif(not ol.have_ref_binary(), reason="oracle/_ref/ref_align not built")
@pytest.mark.parametrize("algo,weights", [
    (ol.LNW, dict(match=3, mismatch=-1, gap_open=-2)), (ol.LNW, dict(match=1, mismatch=-3, gap_open=-1)),
    (ol.LSW, dict(match=3, mismatch=-1, gap_open=-2)), (ol.LSW, dict(match=2, mismatch=-2, gap_open=-1)),
    (ol.ANW, dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1)), (ol.ANW, dict(match=2, mismatch=-1, gap_open=0, gap_extend=-2)),
    (ol.ANW, dict(match=5, mismatch=-4, gap_open=-10, gap_extend=-1)),
])
def test_oracle_matches_live_reference(tmp_path, algo, weights):
    rng = synth.Rng(0xABCD00 + algo * 17 + weights["match"])
    pairs = []
    for k in range(150):
        alpha = [b"0", b"01", b"0123", b"01234"][k % 4]
        R = int(rng.below(1, 70)[0])
        r = synth.random_seq(rng, R, alpha)
        q = synth.mutate(rng, r, 0.1, 0.05, 0.05, alpha) if k % 2 else synth.random_seq(rng, int(rng.below(1, 70)[0]), alpha)
        pairs.append((r, q))
    img = synth.pairs_to_file_bytes(pairs)
    path = tmp_path / "in.txt"
    path.write_bytes(img)
    want = ol.run_reference(algo, str(path), **weights)
    assert oracle_text(algo, img, **weights) == want


def test_threaded_batch_equals_sequential():
    img = open(os.path.join(GOLD, "cfg1_small.in.txt"), "rb").read()
    blob, pairs = ol.parse_image(img)
    p = ol.params(ol.ANW, gap_open=-3, gap_extend=-1)
    a = ol.align_batch(p, blob, pairs, threads=1)
    b = ol.align_batch(p, blob, pairs, threads=4)
    assert (a[0] == b[0]).all() and a[2] == b[2]


# ---- banded pins ------------------------------------------------------------------------------

def test_bsw_full_band_equals_lsw():
    """Pin (i): W >= max(Q,R) => byte-identical to the pinned LSW oracle."""
    for name in ("adversarial", "shapes"):
        img = open(os.path.join(GOLD, f"{name}.in.txt"), "rb").read()
        blob, pairs = ol.parse_image(img)
        W = int(max(pairs["referenceSize"].max(), pairs["querySize"].max()))
        s0, e0, t0 = ol.align_batch(ol.params(ol.LSW), blob, pairs)
        s1, e1, t1 = ol.align_batch(ol.params(ol.BSW, band=W), blob, pairs)
        assert (s0 == s1).all() and (e0 == e1).all() and t0 == t1
        want = open(os.path.join(GOLD, f"{name}.LSW.out.txt"), "rb").read()
        assert ol.format_text(s1, t1) == want


def test_bsw_scores_equal_python_prototype():
    """Pin (ii): scores equal python/LinearBandedSmithWaterman.py with BAND = W+1 (committed fixture)."""
    cases = json.load(open(os.path.join(GOLD, "bsw_python_scores.json")))
    assert len(cases) >= 50
    for c in cases:
        img = synth.pairs_to_file_bytes([(c["ref"].encode(), c["qry"].encode())])
        blob, pairs = ol.parse_image(img)
        s, _, _ = ol.align_batch(ol.params(ol.BSW, band=c["band"]), blob, pairs, strings=False)
        assert int(s[0]) == c["score"], c


def test_bsw_bandmem_and_linear_memory_agree_with_full_matrix():
    """Pin (iii): band-only-memory and rolling-row restatements == full-matrix restatement."""
    rng = synth.Rng(77)
    pairs = []
    for k in range(40):
        R = 50 + int(rng.below(1, 700)[0])
        r = synth.random_seq(rng, R)
        pairs.append((r, synth.mutate(rng, r, 0.05, 0.02, 0.02)))
    blob, idx = ol.parse_image(synth.pairs_to_file_bytes(pairs))
    for W in (0, 1, 7, 64):
        p = ol.params(ol.BSW, band=W)
        a = ol.align_batch(p, blob, idx)
        b = ol.align_batch(p, blob, idx, bandmem=True)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]
        for k, (r, q) in enumerate(pairs[:10]):
            assert ol.lsw_score_only(p, r, q, band=W) == (int(a[0][k]), int(a[1][k][0]), int(a[1][k][1]))
    p = ol.params(ol.LSW)
    a = ol.align_batch(p, blob, idx, strings=False)
    for k, (r, q) in enumerate(pairs):
        assert ol.lsw_score_only(p, r, q) == (int(a[0][k]), int(a[1][k][0]), int(a[1][k][1]))


def test_parse_image_rejects_bad_line_count():
    with pytest.raises(ValueError):
        ol.parse_image(b"0\n0123\n")


@pytest.mark.skipif(not os.path.exists(ol.REF_ALIGN_ALL), reason="reference binary (BACKTRACK_ALL build) not available")
@pytest.mark.parametrize("seed,alphabet,w", [(1, b"01", (3, -1, -2)), (2, b"0", (1, -1, -1)), (3, b"012", (2, -3, -2)), (4, b"0123", (3, -1, -2)),
                                             (5, b"01", (5, -4, -3))])
def test_all_maxima_mode_is_pinned_on_the_reference_built_with_backtrack_all(seed, alphabet, w, tmp_path):
    """c++/LinearSmithWaterman.h:9 BACKTRACK_ALL: one alignment per maximum cell, queued bottom-right first, finished alignments ordered by
    their number of moves.  Byte-identical stdout against the unmodified reference compiled with -DBACKTRACK_ALL, on tie-heavy inputs
    (periodic and low-entropy sequences give many equal maxima).  Pairs that score 0 are left out: there the reference's all-maxima
    mode walks uninitialised cells."""
    m, x, g = w
    rng = synth.Rng(seed)
    pp = []
    while len(pp) < 120:
        R, Q = 1 + int(rng.below(1, 40)[0]), 1 + int(rng.below(1, 40)[0])
        kind = len(pp) % 3
        if kind == 0:
            unit = synth.random_seq(rng, 1 + int(rng.below(1, 4)[0]), alphabet)
            r, q = (unit * 40)[:R], (unit * 40)[1:1 + Q]                    # periodic: many equal maxima
        elif kind == 1:
            r = synth.random_seq(rng, R, alphabet); q = synth.random_seq(rng, Q, alphabet)
        else:
            core = synth.random_seq(rng, 2 + int(rng.below(1, 6)[0]), alphabet)
            r = synth.random_seq(rng, R // 3, alphabet) + core + synth.random_seq(rng, R // 3, alphabet) + core
            q = core + synth.random_seq(rng, Q // 2, alphabet) + core       # repeated core: several copies of the best local alignment
        if set(r) & set(q):                                                 # at least one match: score > 0 with match > 0
            pp.append((r, q))
    img = synth.pairs_to_file_bytes(pp)
    path = tmp_path / "pairs.txt"
    path.write_bytes(bytes(img))
    blob, pairs = ol.parse_image(img)
    txt, n_alignments = ol.lsw_all_text(ol.params(ol.LSW, match=m, mismatch=x, gap_open=g), blob, pairs)
    assert n_alignments > len(pp)                                           # the inputs do have ties
    assert txt == ol.run_reference_all(str(path), m, x, g)
    # the single-path mode's alignment (first maximum in row-major order) is one of them
    s, e, t = ol.align_batch(ol.params(ol.LSW, match=m, mismatch=x, gap_open=g), blob, pairs)
    blocks = txt.split(b" | ")
    assert len(blocks) == len(pp) + 1
    for k in (0, 7, 50, 119):
        assert b"\n".join(t[k]) + b"\n" in blocks[k + 1]


def test_all_maxima_golden_fixture():
    """tests/golden/ties.LSW_ALL.out.txt = stdout of the reference compiled with -DBACKTRACK_ALL (tests/golden/make_golden.py)."""
    img = open(os.path.join(GOLD, "ties.pairs.txt"), "rb").read()
    blob, pairs = ol.parse_image(img)
    txt, n = ol.lsw_all_text(ol.params(ol.LSW), blob, pairs)
    assert txt == open(os.path.join(GOLD, "ties.LSW_ALL.out.txt"), "rb").read()
    assert n > len(pairs)


# ---- affine banded Smith-Waterman (ORC_ABSW): not a reference algorithm, see oracle/dpx_oracle.c:absw_pair -------------------
def _np_local_gotoh_score(r, q, match, mismatch, go, ge, band):
    """Independent restatement (numpy / plain loops, textbook three-state local alignment restricted to |i-j| <= band):
    best score only.  A k-long gap costs go + k*ge, as in the reference's Gotoh (c++/AffineNeedlemanWunsch.cpp)."""
    NEG = -10**9
    Q, R = len(q), len(r)
    H = np.zeros((Q + 1, R + 1), dtype=np.int64)
    D = np.full((Q + 1, R + 1), NEG, dtype=np.int64); I = np.full((Q + 1, R + 1), NEG, dtype=np.int64)
    for i in range(1, Q + 1):
        for j in range(max(1, i - band), min(R, i + band) + 1):
            D[i, j] = max(H[i - 1, j] + go + ge, D[i - 1, j] + ge)
            I[i, j] = max(H[i, j - 1] + go + ge, I[i, j - 1] + ge)
            H[i, j] = max(0, D[i, j], I[i, j], H[i - 1, j - 1] + (match if q[i - 1] == r[j - 1] else mismatch))
    return int(H.max())


@pytest.mark.parametrize("band", [0, 3, 17, 1000])
def test_affine_banded_sw_degenerates_to_the_pinned_linear_band(band):
    """gap_open = 0: a k-long gap costs k * gap_extend and every tie goes to GAP_OPEN, so scores, end cells and strings must be
    byte-identical to the linear banded restatement (itself pinned on the LinearSmithWaterman golden for full bands)."""
    for name in ("adversarial", "shapes", "cfg1_small"):
        blob, pairs = ol.parse_image(open(os.path.join(GOLD, f"{name}.in.txt"), "rb").read())
        for g in (-2, -1):
            a = ol.align_batch(ol.params(ol.ABSW, gap_open=0, gap_extend=g, band=band), blob, pairs)
            b = ol.align_batch(ol.params(ol.BSW, gap_open=g, band=band), blob, pairs)
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]
    # full band, open = 0: the committed LinearSmithWaterman golden text itself
    img = open(os.path.join(GOLD, "mid.in.txt"), "rb").read()
    if band == 1000:
        assert oracle_text(ol.ABSW, img, gap_open=0, gap_extend=-2, band=5000) == open(os.path.join(GOLD, "mid.LSW.out.txt"), "rb").read()


def test_affine_banded_sw_scores_match_an_independent_restatement():
    rng = synth.Rng(0xAB5)
    pp = []
    for k in range(60):
        alpha = [b"01", b"0123"][k % 2]
        r = synth.random_seq(rng, 1 + int(rng.below(1, 60)[0]), alpha)
        q = synth.mutate(rng, r, 0.1, 0.08, 0.08, alpha) if k % 3 else synth.random_seq(rng, 1 + int(rng.below(1, 60)[0]), alpha)
        pp.append((r, q))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for w in (dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1), dict(match=2, mismatch=-3, gap_open=-5, gap_extend=-2), dict(match=1, mismatch=-1, gap_open=-1, gap_extend=0)):
        for band in (2, 9, 100):
            s, e, t = ol.align_batch(ol.params(ol.ABSW, band=band, **w), blob, pairs)
            for k, (r, q) in enumerate(pp):
                assert int(s[k]) == _np_local_gotoh_score(r, q, w["match"], w["mismatch"], w["gap_open"], w["gap_extend"], band)
                # the printed alignment re-scores to the score and spells substrings ending at the end cell
                ref_l, rel, qry_l = t[k]
                if s[k] == 0:
                    assert ref_l == rel == qry_l == b""
                    continue
                sc, in_gap = 0, None
                for a, m, b in zip(ref_l, rel, qry_l):
                    if a == ord("_") or b == ord("_"):
                        kind = "D" if a == ord("_") else "I"
                        sc += w["gap_extend"] + (w["gap_open"] if in_gap != kind else 0)
                        in_gap = kind
                    else:
                        sc += w["match"] if a == b else w["mismatch"]; in_gap = None
                assert sc == int(s[k])
                assert r[: int(e[k][1])].endswith(ref_l.replace(b"_", b"")) and q[: int(e[k][0])].endswith(qry_l.replace(b"_", b""))
