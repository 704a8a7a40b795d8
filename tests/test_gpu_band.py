"""GPU parity of the band-mapped BandedSmithWaterman kernels (csrc/band.cuh: diagonals owned by lanes, int32 scores with
the direction in the two low bits, warp-cooperative backtrack) against the CPU oracle's repaired banded semantics:
scores, end cells and the three alignment strings, bit-exact.  Covers every (M, EXTRA) instantiation (bands 0..96), ragged
and empty pairs, long pairs (several traceback windows, several key blocks) and the fall-back for wider bands."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
KERNEL_BAND, KERNEL_WAVEFRONT = 3, 1


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _check(eng, blob, pairs, band, want_kernel=KERNEL_BAND, **w):
    w = dict(dict(match=3, mismatch=-1, gap_open=-2), **w)
    b = eng.upload(blob, pairs)
    for flags in (ALL, api.OUT_SCORE | api.OUT_END_COORDS):
        b.run(api.make_params(api.BSW, flags=flags, band=band, **w)); b.sync()
        assert b.stats()["kernel_id"] == want_kernel
        res = b.fetch()
        s, e, t = ol.align_batch(ol.params(ol.BSW, band=band, **w), blob, pairs, strings=bool(flags & api.OUT_STRINGS), threads=8)
        bad = np.flatnonzero(res.scores != s)
        assert len(bad) == 0, f"band {band}: score mismatch at {bad[:10]}: gpu {res.scores[bad[:10]]} oracle {s[bad[:10]]}"
        bad = np.flatnonzero((res.end_row_col != e).any(axis=1))
        assert len(bad) == 0, f"band {band}: end-cell mismatch at {bad[:10]}: gpu {res.end_row_col[bad[:5]]} oracle {e[bad[:5]]}"
        if flags & api.OUT_STRINGS:
            for i, (x, y) in enumerate(zip(res.strings, t)):
                assert x == y, f"band {band} pair {i}: strings differ\n gpu {x}\n orc {y}"
    b.free()


def _ragged(seed, n, lo, hi, alphabet=b"0123"):
    rng = synth.Rng(seed)
    pp = []
    for k in range(n):
        R = lo + int(rng.below(1, hi - lo + 1)[0])
        r = synth.random_seq(rng, R, alphabet)
        q = synth.mutate(rng, r, 0.06, 0.03, 0.03, alphabet) if k % 4 else synth.random_seq(rng, lo + int(rng.below(1, hi - lo + 1)[0]), alphabet)
        pp.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp))


@pytest.mark.parametrize("band", [0, 1, 2, 5, 16, 31, 32, 33, 47, 63, 64, 65, 80, 95, 96])
def test_every_band_geometry(eng, band):
    blob, pairs = _ragged(0xB0 + band, 90, 1, 260)
    _check(eng, blob, pairs, band)


@pytest.mark.parametrize("alphabet", [b"0", b"01"])
def test_tie_heavy_alphabets(eng, alphabet):
    blob, pairs = _ragged(3, 80, 1, 150, alphabet)
    for band in (7, 64):
        _check(eng, blob, pairs, band)


def test_other_weights(eng):
    blob, pairs = _ragged(4, 80, 1, 200)
    for w in (dict(match=2, mismatch=-2, gap_open=-1), dict(match=5, mismatch=-4, gap_open=-7), dict(match=1, mismatch=-1, gap_open=-3)):
        _check(eng, blob, pairs, 20, **w)


def test_empty_and_tiny_pairs(eng):
    pp = [(b"", b""), (b"0123", b""), (b"", b"3210"), (b"0", b"0"), (b"0", b"1"), (b"01230123", b"01230123")]
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for band in (0, 3, 64):
        _check(eng, blob, pairs, band)


def test_long_pairs_config4_shape(eng):
    """BASELINE config 4 shape scaled down: long reads, band 64, drift well inside the band; the walk crosses many 64-step
    traceback windows."""
    img = synth.mutated_fixed_file_bytes(5, 6000, 6000, 0x5EED0004, 0.05, 0.01, 0.01)
    blob, pairs = ol.parse_image(img)
    _check(eng, blob, pairs, 64)


def test_config4_true_shape(eng):
    """BASELINE config 4 at its REAL per-pair size: 64 pairs of 10 000 x 10 000 bp, band 64, scores + end cells + strings against
    the band-memory oracle (orc bsw_pair_bandmem: O(Q * W) memory, itself pinned on the full-matrix restatement)."""
    img = synth.mutated_fixed_file_bytes(64, 10_000, 10_000, 0x5EED0004, 0.05, 0.01, 0.01)
    blob, pairs = ol.parse_image(img)
    s, e, t = ol.align_batch(ol.params(ol.BSW, band=64), blob, pairs, strings=True, threads=8, bandmem=True)
    for src in ("raw", "sidecar"):
        if src == "sidecar":
            inp = api.parse_image_native(img); blob, pairs = inp.sequences, inp.pairs
        b = eng.upload(blob, pairs)
        b.run(api.make_params(api.BSW, flags=ALL, band=64)); b.sync()
        assert b.stats()["kernel_id"] == KERNEL_BAND
        res = b.fetch()
        b.free()
        assert (res.scores == s).all() and (res.end_row_col == e).all(), src
        assert res.strings == t, src
        assert min(len(x[0]) for x in res.strings) > 9000                     # the alignments really span the reads


def test_rectangular_pairs_leave_the_band(eng):
    rng = synth.Rng(12)
    pp = [(synth.random_seq(rng, 900), synth.random_seq(rng, 300)), (synth.random_seq(rng, 200), synth.random_seq(rng, 1000))]
    r = synth.random_seq(rng, 700)
    pp.append((r, r[150:650]))          # best local alignment sits on a diagonal 150 off the main one: outside band 64
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for band in (10, 64, 96):
        _check(eng, blob, pairs, band)


def test_wider_bands_and_odd_weights_fall_back(eng):
    blob, pairs = _ragged(5, 40, 1, 200)
    _check(eng, blob, pairs, 97, want_kernel=KERNEL_WAVEFRONT)
    _check(eng, blob, pairs, 20, want_kernel=KERNEL_WAVEFRONT, match=3, mismatch=0, gap_open=-2)


@pytest.mark.parametrize("n", [8192, 13001])
def test_large_batches_with_strings_and_repeated_runs(eng, n):
    """Thousands of pairs with strings (several waves of warps, the traceback slab of the whole batch), run twice on the same batch:
    the second run reuses the slab and the string buffers."""
    blob, pairs = synth.ragged_mutated_blob_pairs(n, 20, 90, 0xBA5D + n, 0.05, 0.02, 0.02)
    b = eng.upload(blob, pairs)
    s, e, t = ol.align_batch(ol.params(ol.BSW, band=9), blob, pairs, strings=True, threads=16)
    for _ in range(2):
        b.run(api.make_params(api.BSW, flags=ALL, band=9)); b.sync()
        assert b.stats()["kernel_id"] == KERNEL_BAND
        res = b.fetch()
        assert (res.scores == s).all() and (res.end_row_col == e).all()
        assert res.strings == t
    b.free()
