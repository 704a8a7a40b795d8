"""GPU parity of the packed two-pair Needleman-Wunsch / Gotoh / Smith-Waterman-with-traceback kernels (csrc/pairwf.cuh: int16x2, directions in the low
score bits) against the CPU oracle: scores and the three alignment strings, bit-exact.  Covers ragged duos, odd pair
counts, empty sequences, several passes (Q > 256), tie-heavy alphabets, weight sets at the edge of the kernel's range,
and the fall-back to the int32 wavefront kernel when a batch does not fit."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
KERNEL_PAIR, KERNEL_PAIR_INT32, KERNEL_WAVEFRONT = 5, 6, 1


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _check(eng, algo, blob, pairs, want_kernel=None, **w):
    b = eng.upload(blob, pairs)
    for flags in (ALL, api.OUT_SCORE | api.OUT_END_COORDS):
        b.run(api.make_params(algo, flags=flags, **w)); b.sync()
        if want_kernel is not None and (algo != api.LSW or flags & api.OUT_STRINGS):      # LSW without strings: short-read kernel
            assert b.stats()["kernel_id"] == want_kernel
        res = b.fetch()
        s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs, strings=bool(flags & api.OUT_STRINGS), threads=8)
        bad = np.flatnonzero(res.scores != s)
        assert len(bad) == 0, f"score mismatch at pairs {bad[:10]}: gpu {res.scores[bad[:10]]} oracle {s[bad[:10]]}"
        assert (res.end_row_col == e).all()
        if flags & api.OUT_STRINGS:
            for i, (x, y) in enumerate(zip(res.strings, t)):
                assert x == y, f"pair {i}: strings differ\n gpu {x}\n orc {y}"
    b.free()


def _ragged(seed, n, lo, hi, alphabet=b"0123", related=True):
    rng = synth.Rng(seed)
    pp = []
    for k in range(n):
        R = lo + int(rng.below(1, hi - lo + 1)[0])
        r = synth.random_seq(rng, R, alphabet)
        if related and k % 3:
            q = synth.mutate(rng, r, 0.05, 0.03, 0.03, alphabet)
        else:
            q = synth.random_seq(rng, lo + int(rng.below(1, hi - lo + 1)[0]), alphabet)
        pp.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp))


WEIGHTS = {
    api.LNW: [dict(match=3, mismatch=-1, gap_open=-2), dict(match=1, mismatch=-1, gap_open=-2), dict(match=5, mismatch=-3, gap_open=-4),
              dict(match=2, mismatch=0, gap_open=-1)],
    api.LSW: [dict(match=3, mismatch=-1, gap_open=-2), dict(match=1, mismatch=-1, gap_open=-2), dict(match=5, mismatch=-3, gap_open=-4),
              dict(match=2, mismatch=-1, gap_open=-2)],
    api.ANW: [dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1), dict(match=2, mismatch=-2, gap_open=-2, gap_extend=-1),
              dict(match=4, mismatch=-1, gap_open=0, gap_extend=-2), dict(match=1, mismatch=-1, gap_open=-4, gap_extend=0)],
}


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
@pytest.mark.parametrize("wi", [0, 1, 2, 3])
def test_ragged_batches(eng, algo, wi):
    blob, pairs = _ragged(100 + wi, 301, 1, 120)
    _check(eng, algo, blob, pairs, KERNEL_PAIR, **WEIGHTS[algo][wi])


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
@pytest.mark.parametrize("alphabet", [b"0", b"01", b"012"])
def test_tie_heavy_alphabets(eng, algo, alphabet):
    blob, pairs = _ragged(7, 200, 1, 90, alphabet, related=False)
    _check(eng, algo, blob, pairs, KERNEL_PAIR, **WEIGHTS[algo][0])


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
def test_multi_pass_queries(eng, algo):
    """Q > 256 rows takes several passes through the shared-memory boundary row; R, Q around the pass and step edges."""
    pp = []
    rng = synth.Rng(11)
    for (R, Q) in [(255, 256), (256, 257), (300, 513), (700, 511), (33, 600), (600, 31), (512, 512), (1, 700), (700, 1)]:
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.04, 0.02, 0.02)[:Q]
        q = q + synth.random_seq(rng, Q - len(q))
        pp.append((r, q))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    _check(eng, algo, blob, pairs, KERNEL_PAIR, **WEIGHTS[algo][0])


def test_gotoh_scores_with_16_and_8_rows_per_lane_agree(eng):
    """Without traceback the Gotoh fill keeps 16 rows per lane for queries above 256 rows (two 512-row passes for 1000 rows instead
    of four 256-row ones; the D / I codes stay uncleaned).  Rows around the pass edges of both geometries, every weight set, against the
    oracle and against the 8-row kernel."""
    pp = []
    rng = synth.Rng(16)
    for (R, Q) in [(300, 257), (400, 511), (200, 512), (513, 513), (90, 1023), (1024, 1025), (700, 1536), (33, 2000)]:
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05, 0.02, 0.02)[:Q]
        pp.append((r, q + synth.random_seq(rng, Q - len(q))))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    flags = api.OUT_SCORE | api.OUT_END_COORDS
    for w in WEIGHTS[api.ANW]:
        s, e, _ = ol.align_batch(ol.params(ol.ANW, **w), blob, pairs, strings=False, threads=8)
        got16 = eng.align_batch(api.make_params(api.ANW, flags=flags, **w), blob, pairs)
        eng.set_option("pairwf_k8", 1)
        try:
            got8 = eng.align_batch(api.make_params(api.ANW, flags=flags, **w), blob, pairs)
        finally:
            eng.set_option("pairwf_k8", 0)
        assert (got16.scores == s).all() and (got8.scores == s).all()
        assert (got16.end_row_col == e).all() and (got8.end_row_col == e).all()


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
def test_config3_shape(eng, algo):
    """BASELINE config 3 shape (1000 x 1000, mutated 2% / 0.5% / 0.5%), a small batch with an odd pair count."""
    img = synth.mutated_fixed_file_bytes(33, 1000, 1000, 0x5EED0003, 0.02, 0.005, 0.005)
    blob, pairs = ol.parse_image(img)
    _check(eng, algo, blob, pairs, KERNEL_PAIR, **WEIGHTS[algo][0])


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
def test_empty_sequences_inside_a_batch(eng, algo):
    pp = [(b"", b""), (b"0123", b""), (b"", b"3210"), (b"0", b"0"), (b"0123012301", b"0123012301"), (b"", b"1")]
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    _check(eng, algo, blob, pairs, KERNEL_PAIR, **WEIGHTS[algo][0])


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
def test_out_of_range_batches_take_the_int32_variant_or_the_wavefront_kernel(eng, algo):
    # scores beyond the int16 budget (4 * (hi - lo) must stay below 2^15): the same kernel in int32, one pair per warp
    blob, pairs = _ragged(5, 7, 2800, 2900)
    _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][0])
    # a mismatch worse than a gap open does not fit the score table: the general wavefront kernel
    blob, pairs = _ragged(6, 50, 1, 80)
    w = dict(match=3, mismatch=-9, gap_open=-2, gap_extend=-1) if algo == api.ANW else dict(match=3, mismatch=-9, gap_open=-2)
    _check(eng, algo, blob, pairs, KERNEL_WAVEFRONT, **w)


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
def test_int32_variant_on_the_small_cases(eng, algo):
    """The pairwf_int32 option forces the one-pair-per-warp int32 instantiation of the same kernels onto inputs the packed one handles."""
    with eng.options(pairwf_int32=1):
        blob, pairs = _ragged(21, 151, 1, 140)
        _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][0])
        pp = [(b"", b""), (b"0123", b""), (b"", b"3210"), (b"0", b"0"), (b"0123012301", b"0123012301")]
        rng = synth.Rng(4)
        r = synth.random_seq(rng, 600); pp.append((r, synth.mutate(rng, r, 0.04, 0.02, 0.02)))
        blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
        _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][1])


@pytest.mark.parametrize("algo", [api.LNW, api.ANW, api.LSW])
@pytest.mark.parametrize("alphabet", [b"01234", b"ACGTNacg"])
def test_five_to_eight_symbol_alphabets_use_the_wide_tables(eng, algo, alphabet):
    """The reference's data sets use '0'..'4'; 5..8 symbols run the int32 instantiation with both table registers for one pair."""
    blob, pairs = _ragged(31, 120, 1, 150, alphabet)
    _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][0])
    a = alphabet
    pp = [(b"", b""), (a[:5], b""), (b"", a[4::-1]), (a[4:5], a[4:5]), (a[4:5], a[0:1])]
    rng = synth.Rng(8)
    r = synth.random_seq(rng, 700, alphabet); pp.append((r, synth.mutate(rng, r, 0.04, 0.02, 0.02, alphabet)))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp + [(alphabet, alphabet[::-1])]))
    _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][2])


def test_nine_symbols_fall_back_to_the_byte_compare_kernel(eng):
    blob, pairs = _ragged(32, 60, 1, 100, b"012345678")
    _check(eng, api.LNW, blob, pairs, KERNEL_WAVEFRONT, **WEIGHTS[api.LNW][0])


def test_very_long_references_fall_back(eng):
    """The per-warp column table of the pair-wavefront kernels lives in shared memory: beyond 12 kbp the wavefront kernel runs."""
    rng = synth.Rng(77)
    r = synth.random_seq(rng, 12500); q = synth.mutate(rng, r[:300], 0.03, 0.01, 0.01)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q), (q, r[:500])]))
    _check(eng, api.LNW, blob, pairs, KERNEL_WAVEFRONT, **WEIGHTS[api.LNW][0])


def test_long_references_inside_the_limit(eng):
    rng = synth.Rng(78)
    r = synth.random_seq(rng, 11000); q = synth.mutate(rng, r[:400], 0.03, 0.01, 0.01)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q), (q, r[:600]), (r[:5000], r[100:4800])]))
    for algo in (api.LNW, api.ANW, api.LSW):
        _check(eng, algo, blob, pairs, KERNEL_PAIR_INT32, **WEIGHTS[algo][0])
