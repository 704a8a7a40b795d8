"""One long pair through the systolic warp chain (csrc/longpair.cuh) against the rolling-row CPU oracle."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu


def _pair(R, Q, seed, sub=0.02, indel=0.005):
    rng = synth.Rng(seed)
    r = synth.random_seq(rng, R)
    if sub is None:
        return r, synth.random_seq(rng, Q)
    q = synth.mutate(rng, r[: int(Q * 1.05)], sub, indel, indel)
    q = (q + synth.random_seq(rng, Q))[:Q]
    return r, q


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("R,Q", [(1, 1), (5, 3), (31, 33), (64, 64), (257, 100), (1000, 1200), (4097, 3000), (20000, 9000), (7000, 30000)])
def test_long_pair_matches_oracle(eng, R, Q):
    for seed, sub in ((1, 0.02), (2, None)):
        r, q = _pair(R, Q, seed * 1000 + R + Q, sub)
        p = api.make_params(api.LSW)
        assert eng.align_long_pair(p, r, q) == ol.lsw_score_only(ol.params(ol.LSW), r, q)


@pytest.mark.parametrize("k,cap", [("2", "8"), ("4", "16"), ("8", "4"), ("16", "4"), ("2", "0"), ("8", "0")])
def test_long_pair_forced_geometry_and_passes(eng, k, cap):
    """Every lane width, several passes over column super-blocks (full-length boundary arrays between passes), ring wrap-around."""
    with eng.options(long_k=int(k), long_cap=int(cap)):
        r, q = _pair(5000, 6000, 77)
        p = api.make_params(api.LSW, match=2, mismatch=-3, gap_open=-2)
        assert eng.align_long_pair(p, r, q) == ol.lsw_score_only(ol.params(ol.LSW, match=2, mismatch=-3, gap_open=-2), r, q)
        # tie-heavy input: the end cell must be the first maximum in row-major order
        r2, q2 = b"01" * 1500, b"10" * 1700
        assert eng.align_long_pair(api.make_params(api.LSW), r2, q2) == ol.lsw_score_only(ol.params(ol.LSW), r2, q2)


def test_long_pair_equals_batch_engine(eng):
    r, q = _pair(3000, 2500, 5)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q)]))
    res = eng.align_batch(api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS), blob, pairs)
    s, row, col = eng.align_long_pair(api.make_params(api.LSW), r, q)
    assert (s, row, col) == (int(res.scores[0]), int(res.end_row_col[0][0]), int(res.end_row_col[0][1]))


def test_long_pair_byte_kernel_for_wide_alphabets(eng):
    """More than four symbols (or the long_notable option) use the byte-compare kernel instead of the score-table kernel."""
    rng = synth.Rng(9)
    r = synth.random_seq(rng, 3000, b"01234"); q = synth.mutate(rng, r, 0.05, 0.01, 0.01, b"01234")[:2800]
    assert eng.align_long_pair(api.make_params(api.LSW), r, q) == ol.lsw_score_only(ol.params(ol.LSW), r, q)
    with eng.options(long_notable=1):
        r, q = _pair(4000, 5000, 31)
        assert eng.align_long_pair(api.make_params(api.LSW), r, q) == ol.lsw_score_only(ol.params(ol.LSW), r, q)


def test_long_pair_letters_and_odd_weights(eng):
    rng = synth.Rng(10)
    r = synth.random_seq(rng, 5000, b"ACGT"); q = synth.mutate(rng, r, 0.03, 0.01, 0.01, b"ACGT")
    for w in (dict(match=1, mismatch=-1, gap_open=-1), dict(match=5, mismatch=-4, gap_open=-7), dict(match=100, mismatch=-100, gap_open=-120)):
        assert eng.align_long_pair(api.make_params(api.LSW, **w), r, q) == ol.lsw_score_only(ol.params(ol.LSW, **w), r, q)


def test_config5_full_size_against_the_cpu_pin(eng):
    """BASELINE config 5 at its REAL size, 1 Mbp x 1 Mbp (1e12 cells), against tests/golden/cfg5_1m.json: the rolling-row CPU
    oracle run once on the same generated pair (tools/pin_cfg5.py).  Both lane widths the model may choose."""
    import hashlib
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg5_1m.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/cfg5_1m.json not generated yet (tools/pin_cfg5.py)")
    g = json.load(open(path))
    R, Q = g["R"], g["Q"]
    img = synth.mutated_fixed_file_bytes(1, R, Q, int(g["seed"], 16), 0.01, 0.001, 0.001)
    ref = img[2:2 + R].tobytes(); qry = img[3 + R:3 + R + Q].tobytes()
    assert hashlib.sha256(ref).hexdigest() == g["ref_sha256"] and hashlib.sha256(qry).hexdigest() == g["qry_sha256"]
    want = (g["score"], g["end_row"], g["end_col"])
    p = api.make_params(api.LSW, **g["weights"])
    assert eng.align_long_pair(p, ref, qry) == want
    with eng.options(long_k=16):
        assert eng.align_long_pair(p, ref, qry) == want
