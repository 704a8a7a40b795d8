"""Alignment strings of one long pair (csrc/longtrace.cuh: checkpointed forward passes + tile-by-tile walk) against the oracle's
full-matrix backtrack, bit-exact; tile geometries forced through the lane width; structural invariants where the oracle's
full matrix would not fit."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _pair(R, Q, seed, sub=0.03, indel=0.01, alphabet=b"0123"):
    rng = synth.Rng(seed)
    r = synth.random_seq(rng, R, alphabet)
    if sub is None:
        return r, synth.random_seq(rng, Q, alphabet)
    q = synth.mutate(rng, r[: int(Q * 1.05)], sub, indel, indel, alphabet)
    q = (q + synth.random_seq(rng, Q, alphabet))[:Q]
    return r, q


def _oracle(r, q, **w):
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q)]))
    s, e, t = ol.align_batch(ol.params(ol.LSW, **w), blob, pairs)
    return (int(s[0]), int(e[0][0]), int(e[0][1])), t[0]


def _check(eng, r, q, **w):
    end, start, lines, st = eng.align_long_pair_strings(api.make_params(api.LSW, **w), r, q)
    want_end, want = _oracle(r, q, **w)
    assert end == want_end
    assert lines == want, f"lines differ (lengths {len(lines[0])} vs {len(want[0])}, tiles {st['tiles']})"
    # the alignment covers ref[start_col:end_col] and qry[start_row:end_row]
    assert lines[0].replace(b"_", b"") == r[start[1]:end[2]] and lines[2].replace(b"_", b"") == q[start[0]:end[1]]
    return st


@pytest.mark.parametrize("R,Q", [(1, 1), (5, 3), (64, 64), (65, 63), (257, 100), (1000, 1200), (4097, 3000), (9000, 12000)])
def test_strings_match_full_matrix_backtrack(eng, R, Q):
    for seed, sub in ((1, 0.03), (2, None)):
        r, q = _pair(R, Q, seed * 1000 + R + Q, sub)
        _check(eng, r, q)


@pytest.mark.parametrize("k", ["2", "4", "8", "16"])
def test_every_tile_geometry_gives_the_same_strings(eng, k):
    """The long_k option forces the lane width, hence the checkpoint spacing: tiles of 64, 128, 256 and 512 rows and columns."""
    with eng.options(long_k=int(k)):
        r, q = _pair(5000, 6000, 77)
        st = _check(eng, r, q, match=2, mismatch=-3, gap_open=-2)
        assert st["tiles"] >= (5000 // (32 * int(k))) // 2
        # gap-rich alignment: long horizontal and vertical runs across tile edges
        rng = synth.Rng(5)
        r = synth.random_seq(rng, 3000)
        q = r[:700] + r[1100:2000] + synth.random_seq(rng, 300) + r[2000:]
        _check(eng, r, q, match=3, mismatch=-4, gap_open=-1)


@pytest.mark.parametrize("tiles_per_round", ["1", "2", "5"])
def test_mispredicted_rounds_only_cost_rounds(eng, tiles_per_round):
    """The tiles ahead of the walk are predicted and filled in bulk; the long_bt_tiles option shrinks a round to a few tiles so that
    the walk keeps stepping onto tiles nobody predicted (and, with long gaps, off the predicted diagonal)."""
    with eng.options(long_k=2, long_bt_tiles=int(tiles_per_round)):
        rng = synth.Rng(15)
        r = synth.random_seq(rng, 3000)
        q = r[:500] + r[900:1500] + synth.random_seq(rng, 400) + r[1500:]
        st = _check(eng, r, q, match=3, mismatch=-4, gap_open=-1)
        assert st["rounds"] >= st["tiles"] // int(tiles_per_round)
        r, q = _pair(2500, 2600, 16)
        _check(eng, r, q)


def test_several_passes_with_checkpoints(eng):
    with eng.options(long_k=2, long_cap=8):
        r, q = _pair(4000, 3500, 12)
        _check(eng, r, q)


def test_ties_wide_alphabets_and_zero_score(eng):
    _check(eng, b"01" * 1500, b"10" * 1700)
    _check(eng, b"0" * 2000, b"0" * 1500)
    rng = synth.Rng(9)
    r = synth.random_seq(rng, 3000, b"ACGTN"); q = synth.mutate(rng, r, 0.05, 0.01, 0.01, b"ACGTN")[:2800]
    _check(eng, r, q)
    end, start, lines, st = eng.align_long_pair_strings(api.make_params(api.LSW), b"0" * 500, b"1" * 400)
    assert end == (0, 0, 0) and lines == (b"", b"", b"")


def test_strings_equal_the_batch_engine(eng):
    r, q = _pair(3000, 2500, 5)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q)]))
    res = eng.align_batch(api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS), blob, pairs)
    end, start, lines, st = eng.align_long_pair_strings(api.make_params(api.LSW), r, q)
    assert lines == res.strings[0] and end[0] == int(res.scores[0])


def alignment_score(lines, m, x, g):
    ref, rel, qry = (np.frombuffer(s, dtype=np.uint8) for s in lines)
    gaps = int((rel == ord(" ")).sum())
    return m * int((rel == ord("*")).sum()) + x * int((rel == ord("|")).sum()) + g * gaps


def test_200kbp_pair_invariants(eng):
    """Beyond the oracle's full matrix: the printed alignment must re-score to the forward pass's score, spell the two
    subsequences it claims to cover, and start in a cell the walk could stop at."""
    r, q = _pair(200_000, 200_000, 4242, sub=0.01, indel=0.001)
    end, start, lines, st = eng.align_long_pair_strings(api.make_params(api.LSW), r, q)
    assert end == ol.lsw_score_only(ol.params(ol.LSW), r, q)
    assert alignment_score(lines, 3, -1, -2) == end[0]
    assert lines[0].replace(b"_", b"") == r[start[1]:end[2]] and lines[2].replace(b"_", b"") == q[start[0]:end[1]]
    assert len(lines[0]) == len(lines[1]) == len(lines[2])
    assert st["tiles"] <= 200_000 // st["tile_rows"] + 200_000 // st["tile_cols"] + 2
