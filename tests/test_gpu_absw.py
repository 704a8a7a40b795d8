"""GPU parity of the affine banded Smith-Waterman (DPX_ALGO_ABSW; not a reference algorithm — include/dpxalign.h, oracle/dpx_oracle.c:
absw_pair) against the CPU restatement: scores, end cells, strings, bit-exact; plus the degenerate case gap_open = 0, which must
reproduce the linear BandedSmithWaterman (specialised band kernel) byte for byte."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _ragged(seed, n, lo, hi, alphabet=b"0123"):
    rng = synth.Rng(seed)
    pp = [(b"", b""), (b"0123", b""), (b"", b"3210"), (b"0", b"0"), (b"0", b"1")]
    for k in range(n):
        r = synth.random_seq(rng, lo + int(rng.below(1, hi - lo + 1)[0]), alphabet)
        q = synth.mutate(rng, r, 0.06, 0.04, 0.04, alphabet) if k % 4 else synth.random_seq(rng, lo + int(rng.below(1, hi - lo + 1)[0]), alphabet)
        pp.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp))


def _check(eng, blob, pairs, band, **w):
    s, e, t = ol.align_batch(ol.params(ol.ABSW, band=band, **w), blob, pairs, threads=8)
    res = eng.align_batch(api.make_params(api.ABSW, flags=ALL, band=band, **w), blob, pairs)
    bad = np.flatnonzero(res.scores != s)
    assert len(bad) == 0, f"band {band} {w}: scores differ at {bad[:8]}: gpu {res.scores[bad[:8]]} oracle {s[bad[:8]]}"
    bad = np.flatnonzero((res.end_row_col != e).any(axis=1))
    assert len(bad) == 0, f"band {band} {w}: end cells differ at {bad[:8]}"
    for i, (x, y) in enumerate(zip(res.strings, t)):
        assert x == y, f"band {band} {w} pair {i}:\n gpu {x}\n orc {y}"
    res2 = eng.align_batch(api.make_params(api.ABSW, flags=api.OUT_SCORE | api.OUT_END_COORDS, band=band, **w), blob, pairs)
    assert (res2.scores == s).all() and (res2.end_row_col == e).all()


@pytest.mark.parametrize("band", [0, 1, 5, 31, 64, 200, 5000])
def test_bands(eng, band):
    blob, pairs = _ragged(7 + band, 70, 1, 400)
    _check(eng, blob, pairs, band, match=3, mismatch=-1, gap_open=-3, gap_extend=-1)


@pytest.mark.parametrize("w", [dict(match=2, mismatch=-3, gap_open=-5, gap_extend=-2), dict(match=1, mismatch=-1, gap_open=-1, gap_extend=0),
                               dict(match=5, mismatch=-4, gap_open=-10, gap_extend=-1), dict(match=3, mismatch=-1, gap_open=0, gap_extend=-2)])
def test_weights_and_alphabets(eng, w):
    for alphabet in (b"0", b"01", b"01234", b"ACGTN"):
        blob, pairs = _ragged(3, 40, 1, 150, alphabet)
        for band in (4, 40):
            _check(eng, blob, pairs, band, **w)


def test_long_queries_cross_stripes(eng):
    """Queries longer than one 256-row stripe: D and H cross stripes through the per-warp boundary rows, windows start past column 1."""
    img = synth.mutated_fixed_file_bytes(6, 1500, 1400, 0xAB5, 0.05, 0.02, 0.02)
    blob, pairs = ol.parse_image(img)
    for band in (20, 64, 3000):
        _check(eng, blob, pairs, band, match=3, mismatch=-1, gap_open=-3, gap_extend=-1)


def test_zero_open_cost_is_the_linear_banded_aligner(eng):
    blob, pairs = _ragged(11, 120, 1, 300)
    for band in (6, 64):
        a = eng.align_batch(api.make_params(api.ABSW, flags=ALL, band=band, gap_open=0, gap_extend=-2), blob, pairs)
        b = eng.align_batch(api.make_params(api.BSW, flags=ALL, band=band, gap_open=-2), blob, pairs)
        assert (a.scores == b.scores).all() and (a.end_row_col == b.end_row_col).all() and a.strings == b.strings


def test_class_and_text_and_sidecar(eng, capsysbinary):
    rng = synth.Rng(5)
    r = synth.random_seq(rng, 300); q = synth.mutate(rng, r, 0.05, 0.03, 0.03)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q)]))
    s, e, t = ol.align_batch(ol.params(ol.ABSW, band=16, gap_open=-3, gap_extend=-1), blob, pairs)
    api.AffineBandedSmithWaterman(r, q, 3, -1, -3, -1, 9, band_width=16).align()
    assert capsysbinary.readouterr().out == ol.format_text(s, t, 9)
    inp = api.parse_image_native(synth.pairs_to_file_bytes([(r, q)] * 50))
    txt = eng.align_batch_text(api.make_params(api.ABSW, flags=ALL, band=16, gap_open=-3, gap_extend=-1), inp.sequences, inp.pairs)
    assert txt == ol.format_text(np.repeat(s, 50), t * 50)
    inp.free()
