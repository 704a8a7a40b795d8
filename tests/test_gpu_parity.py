"""GPU parity (the tests proper): libdpxalign through its C ABI against
  (1) the golden fixtures produced by the compiled, unmodified reference classes (byte-identical text),
  (2) the CPU oracle on seeded random / adversarial inputs (scores, end cells, strings),
  (3) the reference's FakeDPX known-answer vectors replayed on the hardware DPX instructions.
Integer work: the bar is bit-exact."""
import glob
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(GOLD, "*.in.txt")))
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
KW = {api.LNW: dict(gap_open=-2), api.LSW: dict(gap_open=-2), api.ANW: dict(gap_open=-3, gap_extend=-1)}
NAMES = {api.LNW: "LNW", api.ANW: "ANW", api.LSW: "LSW", api.BSW: "BSW"}


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def gpu_vs_oracle(eng, algo, blob, pairs, **w):
    res = eng.align_batch(api.make_params(algo, flags=ALL, **w), blob, pairs)
    s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs)
    bad = np.flatnonzero(res.scores != s)
    assert len(bad) == 0, f"score mismatch at pairs {bad[:10]}: gpu {res.scores[bad[:10]]} oracle {s[bad[:10]]}"
    if algo in (api.LSW, api.BSW):
        bad = np.flatnonzero((res.end_row_col != e).any(axis=1))
        assert len(bad) == 0, f"end-cell mismatch at {bad[:10]}: gpu {res.end_row_col[bad[:5]]} oracle {e[bad[:5]]}"
    for i, (a, b) in enumerate(zip(res.strings, t)):
        assert a == b, f"pair {i}: strings differ\n gpu {a}\n orc {b}"
    # score-only mode must give the same scores / end cells
    res2 = eng.align_batch(api.make_params(algo, flags=api.OUT_SCORE | api.OUT_END_COORDS, **w), blob, pairs)
    assert (res2.scores == s).all()
    if algo in (api.LSW, api.BSW):
        assert (res2.end_row_col == e).all()


@pytest.mark.parametrize("name", SETS)
@pytest.mark.parametrize("algo", [api.LNW, api.LSW, api.ANW])
def test_golden_text_is_byte_identical(eng, name, algo):
    p = api.parse_input(os.path.join(GOLD, f"{name}.in.txt"))
    res = eng.align_batch(api.make_params(algo, flags=ALL, **KW[algo]), p.sequences, p.pairs)
    want = open(os.path.join(GOLD, f"{name}.{NAMES[algo]}.out.txt"), "rb").read()
    assert res.text() == want


def random_pairs(seed, n, maxlen, alphabets=(b"0", b"01", b"0123", b"01234")):
    rng = synth.Rng(seed)
    pairs = []
    for k in range(n):
        alpha = alphabets[k % len(alphabets)]
        R = int(rng.below(1, maxlen + 1)[0])
        r = synth.random_seq(rng, R, alpha)
        q = synth.mutate(rng, r, 0.08, 0.03, 0.03, alpha) if k % 2 else synth.random_seq(rng, int(rng.below(1, maxlen + 1)[0]), alpha)
        pairs.append((r, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pairs))


@pytest.mark.parametrize("algo,w", [
    (api.LNW, dict(match=3, mismatch=-1, gap_open=-2)), (api.LNW, dict(match=1, mismatch=-3, gap_open=-1)),
    (api.LSW, dict(match=3, mismatch=-1, gap_open=-2)), (api.LSW, dict(match=2, mismatch=-2, gap_open=-1)),
    (api.ANW, dict(match=3, mismatch=-1, gap_open=-3, gap_extend=-1)), (api.ANW, dict(match=2, mismatch=-1, gap_open=0, gap_extend=-2)),
    (api.ANW, dict(match=5, mismatch=-4, gap_open=-10, gap_extend=-1)),
])
def test_random_pairs_vs_oracle(eng, algo, w):
    blob, pairs = random_pairs(0xC0FFEE + algo, 300, 90)
    gpu_vs_oracle(eng, algo, blob, pairs, **w)


@pytest.mark.parametrize("algo", [api.LNW, api.LSW, api.ANW])
def test_multi_stripe_lengths_vs_oracle(eng, algo):
    """Lengths that cross several 128/256-row stripes and are not multiples of 32."""
    blob, pairs = random_pairs(0xBEEF + algo, 24, 700, alphabets=(b"0123", b"01"))
    gpu_vs_oracle(eng, algo, blob, pairs, **KW[algo])


@pytest.mark.parametrize("band", [0, 1, 3, 17, 64, 1000])
def test_banded_vs_oracle(eng, band):
    blob, pairs = random_pairs(0xBA9D + band, 120, 200, alphabets=(b"0123", b"01", b"0"))
    gpu_vs_oracle(eng, api.BSW, blob, pairs, match=3, mismatch=-1, gap_open=-2, band=band)


@pytest.mark.parametrize("band", [0, 1, 5, 40])
def test_banded_multi_stripe_queries_on_the_byte_kernel(eng, band):
    """Queries of several stripes on the general wavefront kernel (8 symbols keep the band kernel out): the cell above-left of a
    later stripe's first cell lies on the band's edge and belongs to the previous stripe's last row (found by tools/fuzz_gpu.py)."""
    blob, pairs = random_pairs(0xED9E + band, 40, 700, alphabets=(b"ACGTNacg",))
    gpu_vs_oracle(eng, api.BSW, blob, pairs, match=2, mismatch=-5, gap_open=-5, band=band)


def test_long_reference_short_query_smith_waterman_scores(eng):
    """A 3 kbp reference with short queries passes the short-read kernel's score-range test but not its shared-memory budget:
    the batch must fall through to the next kernel instead of failing the launch (found by tools/fuzz_gpu.py)."""
    rng = synth.Rng(99)
    r = synth.random_seq(rng, 3100)
    pp = [(r, r[500:900]), (r[:2600], synth.mutate(rng, r[100:400], 0.05, 0.02, 0.02)), (r[:2500], b"")]
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for flags in (api.OUT_SCORE, api.OUT_SCORE | api.OUT_END_COORDS):
        res = eng.align_batch(api.make_params(api.LSW, match=1, mismatch=-2, gap_open=-3, flags=flags), blob, pairs)
        s, e, _ = ol.align_batch(ol.params(api.LSW, match=1, mismatch=-2, gap_open=-3), blob, pairs, strings=False)
        assert (res.scores == s).all()
        if flags & api.OUT_END_COORDS:
            assert (res.end_row_col == e).all()


def test_banded_full_band_equals_lsw_golden(eng):
    p = api.parse_input(os.path.join(GOLD, "shapes.in.txt"))
    res = eng.align_batch(api.make_params(api.BSW, band=600, flags=ALL), p.sequences, p.pairs)
    assert res.text() == open(os.path.join(GOLD, "shapes.LSW.out.txt"), "rb").read()


def test_banded_python_prototype_scores(eng):
    for c in json.load(open(os.path.join(GOLD, "bsw_python_scores.json"))):
        blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(c["ref"].encode(), c["qry"].encode())]))
        res = eng.align_batch(api.make_params(api.BSW, band=c["band"], flags=api.OUT_SCORE), blob, pairs)
        assert int(res.scores[0]) == c["score"]


def test_empty_batch_and_empty_sequences(eng):
    res = eng.align_batch(api.make_params(api.LNW, flags=ALL), np.zeros(0, np.uint8), np.zeros(0, api.PAIR_DTYPE))
    assert len(res.scores) == 0 and res.strings == []
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(b"", b""), (b"0123", b""), (b"", b"3210")]))
    for algo in (api.LNW, api.LSW, api.ANW):
        gpu_vs_oracle(eng, algo, blob, pairs, **KW[algo])


def test_reference_style_classes_print_reference_bytes(eng, capsysbinary):
    p = api.parse_input(os.path.join(GOLD, "adversarial.in.txt"))
    want = open(os.path.join(GOLD, "adversarial.LNW.out.txt"), "rb").read()
    for i in range(12):                          # c++/main.cpp:243-246
        pr = p.pairs[i]
        r = p.sequences[pr["referenceIdx"]: pr["referenceIdx"] + pr["referenceSize"]].tobytes()
        q = p.sequences[pr["queryIdx"]: pr["queryIdx"] + pr["querySize"]].tobytes()
        api.LinearNeedlemanWunsch(r, q, i, 3, -1, -2).align()
    out = capsysbinary.readouterr().out
    assert want.startswith(out) and out.count(b"\n") == 48


# ---- DPX instruction semantics on hardware ----------------------------------------------------------
OPS = ["__vimax3_s32", "__vimax3_s16x2", "__vimax3_u32", "__vimax3_u16x2", "__vimin3_s32", "__vimin3_s16x2", "__vimin3_u32",
       "__vimin3_u16x2", "__vimax_s32_relu", "__vimax_s16x2_relu", "__vimin_s32_relu", "__vimin_s16x2_relu",
       "__vimax3_s32_relu", "__vimax3_s16x2_relu", "__vimin3_s32_relu", "__vimin3_s16x2_relu",
       "__vibmax_s32", "__vibmax_u32", "__vibmin_s32", "__vibmin_u32", "__vibmax_s16x2", "__vibmax_u16x2",
       "__vibmin_s16x2", "__vibmin_u16x2", "__viaddmax_s32", "__viaddmax_u32", "__viaddmin_s32", "__viaddmin_u32",
       "__viaddmax_s16x2", "__viaddmax_u16x2", "__viaddmin_s16x2", "__viaddmin_u16x2",
       "__viaddmax_s32_relu", "__viaddmin_s32_relu", "__viaddmax_s16x2_relu", "__viaddmin_s16x2_relu"]


def test_fakedpx_known_answers_on_hardware(eng):
    vec = json.load(open(os.path.join(GOLD, "fakedpx_vectors.json")))
    assert len(vec) >= 70
    assert eng.selftest() == 0
    for v in vec:
        op = OPS.index(v["fn"])
        args = [x & 0xFFFFFFFF for x in v["args"]] + [0] * (3 - len(v["args"]))
        out, ph, pl = eng.dpx_eval(op, [args[0]], [args[1]], [args[2]])
        assert int(out[0]) == v["want"] & 0xFFFFFFFF, v
        if "pred_hi" in v["preds"]:
            assert bool(ph[0]) == v["preds"]["pred_hi"], v
        if "pred_low" in v["preds"]:
            assert bool(pl[0]) == v["preds"]["pred_low"], v


def _s16(x):
    x = np.asarray(x, dtype=np.uint32)
    lo = (x & 0xFFFF).astype(np.int64); hi = (x >> 16).astype(np.int64)
    return np.where(lo >= 32768, lo - 65536, lo), np.where(hi >= 32768, hi - 65536, hi)


def test_s16x2_ops_random_vs_numpy(eng):
    """The reference's FakeDPX never tests the s16x2 add/relu variants (c++/testFakeDPX.cpp:108-113 stop at u32);
    check the ones our kernels use against their definition (per-half, wrap-around add)."""
    rng = np.random.default_rng(5)
    n = 4096
    a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    c = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    (al, ah), (bl, bh), (cl, ch) = _s16(a), _s16(b), _s16(c)

    def wrap(x):
        return ((x + 32768) % 65536) - 32768

    def pack(lo, hi):
        return ((hi.astype(np.int64) & 0xFFFF) << 16 | (lo.astype(np.int64) & 0xFFFF)).astype(np.uint32)

    out, _, _ = eng.dpx_eval(OPS.index("__viaddmax_s16x2"), a, b, c)
    assert (out == pack(np.maximum(wrap(al + bl), cl), np.maximum(wrap(ah + bh), ch))).all()
    out, _, _ = eng.dpx_eval(OPS.index("__viaddmax_s16x2_relu"), a, b, c)
    assert (out == pack(np.maximum(np.maximum(wrap(al + bl), cl), 0), np.maximum(np.maximum(wrap(ah + bh), ch), 0))).all()
    out, _, _ = eng.dpx_eval(OPS.index("__vimax3_s16x2_relu"), a, b, c)
    assert (out == pack(np.maximum(np.maximum(np.maximum(al, bl), cl), 0), np.maximum(np.maximum(np.maximum(ah, bh), ch), 0))).all()
    out, ph, pl = eng.dpx_eval(OPS.index("__vibmax_s16x2"), a, b, c)
    assert (out == pack(np.maximum(al, bl), np.maximum(ah, bh))).all()
    assert (ph.astype(bool) == (ah >= bh)).all() and (pl.astype(bool) == (al >= bl)).all()


# ---- packed int16x2 short-read kernel (score + end cell) ------------------------------------------------
def _staged(eng, blob, pairs, params):
    b = eng.upload(blob, pairs)
    b.run(params)
    res = b.fetch()
    st = b.stats()
    b.free()
    return res, st


def _short_eligible(R, Q, w):
    """Mirror of short_eligible() in csrc/dpxalign.cu: unsigned 16-bit keys with k >= 3 position bits, <= 255 step blocks."""
    top = w["match"] * min(R, Q) + max(2, -w["gap_open"])
    k = 0
    while k < 8 and (top << (k + 1)) < 65536:
        k += 1
    return k >= 3 and R <= 4096 and ((R + 16) >> (k - 1)) < 255


@pytest.mark.parametrize("shape", [(150, 150, 4096), (150, 150, 4097), (100, 100, 513), (64, 40, 300), (151, 152, 257),
                                   (250, 250, 128), (300, 100, 64), (33, 400, 64), (1, 1, 70), (500, 160, 40),
                                   (700, 700, 40), (1200, 900, 16)])
@pytest.mark.parametrize("w", [dict(match=3, mismatch=-1, gap_open=-2), dict(match=2, mismatch=-2, gap_open=-1),
                               dict(match=5, mismatch=-4, gap_open=-3)])
def test_short_read_kernel_uniform(eng, shape, w):
    R, Q, n = shape
    img = synth.uniform_file_bytes(n, R, Q, 0x5EED0002 + R * 7 + Q)
    blob, pairs = ol.parse_image(img)
    # make a quarter of the pairs similar so that scores are large and ties / long diagonals occur
    rng = np.random.default_rng(R + Q)
    for k in range(0, n, 4):
        r0, q0 = pairs[k]["referenceIdx"], pairs[k]["queryIdx"]
        m = min(R, Q)
        blob[q0:q0 + m] = blob[r0:r0 + m]
        flips = rng.integers(0, m, max(1, m // 20))
        blob[q0 + flips] = ord("0")
    for flags in (api.OUT_SCORE | api.OUT_END_COORDS, api.OUT_SCORE):
        res, st = _staged(eng, blob, pairs, api.make_params(api.LSW, flags=flags, **w))
        assert st["kernel_id"] == (2 if _short_eligible(R, Q, w) else 1), "unexpected kernel family"
        s, e, _ = ol.align_batch(ol.params(ol.LSW, **w), blob, pairs, strings=False, threads=8)
        assert (res.scores == s).all()
        if flags & api.OUT_END_COORDS:
            bad = np.flatnonzero((res.end_row_col != e).any(axis=1))
            assert len(bad) == 0, f"{len(bad)} end cells differ, first {bad[:5]}: gpu {res.end_row_col[bad[:3]]} oracle {e[bad[:3]]}"


@pytest.mark.parametrize("alphabet", [b"01234", b"ACGTN", b"01234567", b"ACGTNRYK"])
def test_short_read_kernel_wide_alphabets(eng, alphabet):
    """5..8 symbols (the reference's data sets use '0'..'4'): the WIDE instantiation of the short-read kernel -- one pair per lane
    group, byte codes, all eight table bytes of a row for that pair -- instead of the byte-compare wavefront kernel."""
    rng = synth.Rng(len(alphabet) * 101 + alphabet[0])
    pp = [(b"", b""), (alphabet[:1], alphabet[-1:]), (alphabet, alphabet), (alphabet[::-1] * 3, alphabet * 2)]
    for k in range(900):
        R = 1 + int(rng.below(1, 330)[0])
        r = synth.random_seq(rng, R, alphabet)
        q = synth.mutate(rng, r, 0.05, 0.03, 0.03, alphabet) if k % 3 else synth.random_seq(rng, 1 + int(rng.below(1, 330)[0]), alphabet)
        pp.append((r, q))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for w in (dict(match=3, mismatch=-1, gap_open=-2), dict(match=2, mismatch=-3, gap_open=-1), dict(match=5, mismatch=-4, gap_open=-7)):
        for flags in (api.OUT_SCORE | api.OUT_END_COORDS, api.OUT_SCORE):
            res, st = _staged(eng, blob, pairs, api.make_params(api.LSW, flags=flags, **w))
            assert st["kernel_id"] == 2, "5..8 symbols must stay on the short-read kernel"
            s, e, _ = ol.align_batch(ol.params(ol.LSW, **w), blob, pairs, strings=False, threads=8)
            assert (res.scores == s).all()
            if flags & api.OUT_END_COORDS:
                assert (res.end_row_col == e).all()
    # uniform 150 x 150 (config 2's shape) over '0'..'4', several passes for longer queries
    for R, Q, n in ((150, 150, 3000), (400, 333, 200)):
        img = synth.uniform_file_bytes(n, R, Q, 99 + R, alphabet)
        blob, pairs = ol.parse_image(img)
        blob[pairs["queryIdx"][0]: pairs["queryIdx"][0] + min(R, Q)] = blob[pairs["referenceIdx"][0]: pairs["referenceIdx"][0] + min(R, Q)]
        res, st = _staged(eng, blob, pairs, api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS))
        assert st["kernel_id"] == 2
        s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False, threads=8)
        assert (res.scores == s).all() and (res.end_row_col == e).all()
    # a ninth symbol leaves the table kernels
    nine = ol.parse_image(synth.pairs_to_file_bytes([(b"012345678", b"876543210"), (b"0123", b"0123")]))
    res, st = _staged(eng, nine[0], nine[1], api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS))
    assert st["kernel_id"] == 1
    s, e, _ = ol.align_batch(ol.params(ol.LSW), nine[0], nine[1], strings=False)
    assert (res.scores == s).all() and (res.end_row_col == e).all()


def test_short_read_kernel_ragged_lengths(eng):
    rng = synth.Rng(0x5EED0000 + 77)
    pairs = []
    for k in range(1501):
        R = 1 + int(rng.below(1, 300)[0])
        r = synth.random_seq(rng, R, b"0123" if k % 3 else b"01")
        q = synth.mutate(rng, r, 0.05, 0.02, 0.02) if k % 2 else synth.random_seq(rng, 1 + int(rng.below(1, 300)[0]))
        pairs.append((r, q))
    pairs += [(b"", b"0123"), (b"0123", b""), (b"", b"")]
    blob, idx = ol.parse_image(synth.pairs_to_file_bytes(pairs))
    res, st = _staged(eng, blob, idx, api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS))
    assert st["kernel_id"] == 2
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, idx, strings=False, threads=8)
    assert (res.scores == s).all() and (res.end_row_col == e).all()


def test_five_symbol_alphabet_uses_the_wide_tables(eng):
    """'0'..'4' (the reference's data sets): score requests stay on the short-read kernel (WIDE); with the kernel switched off
    the byte-compare wavefront kernel gives the same bytes."""
    blob, idx = random_pairs(0x44, 200, 120, alphabets=(b"01234",))
    with eng.options(no_shortread=1):
        res0, st0 = _staged(eng, blob, idx, api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS))
    assert st0["kernel_id"] == 1
    res, st = _staged(eng, blob, idx, api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS))
    assert st["kernel_id"] == 2
    assert (res0.scores == res.scores).all() and (res0.end_row_col == res.end_row_col).all()
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, idx, strings=False)
    assert (res.scores == s).all() and (res.end_row_col == e).all()


def test_one_call_pipeline_matches_staged_path(eng):
    """dpx_align_batch cuts large score/end-cell batches into chunks on several streams; results must not depend on it."""
    n = 300_000
    blob, pairs = synth.uniform_blob_pairs(n, 60, 50, 0x5EED0000 + 9)
    p = api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS)
    one = eng.align_batch(p, blob, pairs)
    staged, st = _staged(eng, blob, pairs, p)
    assert (one.scores == staged.scores).all() and (one.end_row_col == staged.end_row_col).all()
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs[:20000], strings=False, threads=8)
    assert (one.scores[:20000] == s).all() and (one.end_row_col[:20000] == e).all()
    # ragged lengths through the same pipeline (per-chunk sort + scan)
    rng = np.random.default_rng(3)
    pairs2 = pairs.copy()
    pairs2["referenceSize"] = rng.integers(0, 61, n)
    pairs2["querySize"] = rng.integers(0, 51, n)
    one = eng.align_batch(p, blob, pairs2)
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs2[-20000:], strings=False, threads=8)
    assert (one.scores[-20000:] == s).all() and (one.end_row_col[-20000:] == e).all()


def test_invalid_index_is_rejected(eng):
    blob, pairs = synth.uniform_blob_pairs(10, 20, 20, 1)
    bad = pairs.copy()
    bad["referenceIdx"][3] = len(blob) - 5
    with pytest.raises(api.DpxError) as e:
        eng.align_batch(api.make_params(api.LSW), blob, bad)
    assert e.value.status == -1


@pytest.mark.parametrize("late_symbol", [ord("3"), ord("4")])
def test_one_call_pipeline_redoes_when_alphabet_grows(eng, late_symbol):
    """Chunk 0 fixes the alphabet map of the pipelined call; a symbol that first appears in a later chunk must not
    change any result (the pack kernel flags it and the call is redone through the single-batch route)."""
    n = 300_000
    blob, pairs = synth.uniform_blob_pairs(n, 40, 36, 0x5EED0000 + 21, alphabet=b"012")
    blob = blob.copy()
    tail = pairs[-5000:]
    rng = np.random.default_rng(4)
    for k in range(0, 5000, 7):
        blob[tail["referenceIdx"][k] + int(rng.integers(0, 40))] = late_symbol
        blob[tail["queryIdx"][k] + int(rng.integers(0, 36))] = late_symbol
    p = api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS)
    one = eng.align_batch(p, blob, pairs)
    for sl in (slice(0, 3000), slice(n - 6000, n)):
        s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs[sl], strings=False, threads=8)
        assert (one.scores[sl] == s).all() and (one.end_row_col[sl] == e).all()


@pytest.mark.parametrize("algo", [api.LNW, api.LSW, api.ANW, api.BSW])
def test_gpu_formatted_text_equals_host_formatting(eng, algo):
    """dpx_align_batch_text: the reference's stdout blocks written by the GPU (digits, separators, strings, empty SW lines)."""
    blob, pairs = random_pairs(0x7E87 + algo, 333, 120)
    w = dict(KW.get(algo, dict(gap_open=-2)))
    if algo == api.BSW:
        w["band"] = 9
    res = eng.align_batch(api.make_params(algo, flags=ALL, **w), blob, pairs)
    assert eng.align_batch_text(api.make_params(algo, flags=ALL, **w), blob, pairs) == res.text()
    assert eng.align_batch_text(api.make_params(algo, flags=ALL, **w), blob, pairs, first_index=99990) == res.text(99990)
    lines = b"".join(b"%d | %d\n" % (i, int(s)) for i, s in enumerate(res.scores))
    assert eng.align_batch_text(api.make_params(algo, flags=api.OUT_SCORE, **w), blob, pairs) == lines


def test_gpu_formatted_text_golden(eng):
    for algo in (api.LNW, api.LSW, api.ANW):
        p = api.parse_input(os.path.join(GOLD, "adversarial.in.txt"))
        want = open(os.path.join(GOLD, f"adversarial.{NAMES[algo]}.out.txt"), "rb").read()
        assert eng.align_batch_text(api.make_params(algo, flags=ALL, **KW[algo]), p.sequences, p.pairs) == want
    assert eng.align_batch_text(api.make_params(api.LNW, flags=ALL), np.zeros(0, np.uint8), np.zeros(0, api.PAIR_DTYPE)) == b""


def test_device_parser_matches_host_parser(eng, tmp_path):
    """dpx_batch_upload_image / dpx_align_file_text: newline scan + seqPair index + pack on the GPU give the same batch as
    parseInput + dpx_batch_upload (c++/parseInput.cpp:78-113), including inputInfo and the format error."""
    for name in ("adversarial", "cfg1_small", "mid", "shapes"):
        path = os.path.join(GOLD, f"{name}.in.txt")
        img = open(path, "rb").read()
        p = api.parse_input(path)
        for algo in (api.LNW, api.LSW, api.ANW):
            want = eng.align_batch(api.make_params(algo, flags=ALL, **KW[algo]), p.sequences, p.pairs).text()
            b = eng.upload_image(img)
            assert b.n == len(p.pairs)
            for k in ("numPairs", "numBytes", "numCells", "minReferenceLength", "maxReferenceLength", "minQueryLength", "maxQueryLength"):
                assert b.info[k] == p.info[k], k
            assert abs(b.info["avgReferenceLength"] - p.info["avgReferenceLength"]) < 1e-9
            b.run(api.make_params(algo, flags=ALL, **KW[algo])); b.sync()
            assert b.fetch().text() == want
            b.free()
            text, info = eng.align_file_text(api.make_params(algo, flags=ALL, **KW[algo]), path)
            assert text == want and info["numPairs"] == len(p.pairs)
    bad = tmp_path / "bad.txt"
    bad.write_bytes(b"0\n0123\n0123\n1\n0123\n")
    with pytest.raises(api.DpxError) as e:
        eng.align_file_text(api.make_params(api.LNW, flags=ALL), str(bad))
    assert e.value.status == -6
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    assert eng.align_file_text(api.make_params(api.LNW, flags=ALL), str(empty))[0] == b""


def test_fasta_pairs_align_like_the_same_pairs_in_the_reference_format(eng, tmp_path):
    """FASTA records through dpx_parse_fastx give the blob + index every entry point takes."""
    rng = synth.Rng(41)
    pp = []
    for k in range(40):
        r = synth.random_seq(rng, 80 + 5 * k, b"ACGT")
        pp.append((r, synth.mutate(rng, r, 0.06, 0.03, 0.03, b"ACGT")))
    fa = tmp_path / "pairs.fa"
    fa.write_bytes(b"".join(b">r%d\n" % k + b"\n".join(r[i:i + 60] for i in range(0, len(r), 60)) + b"\n>q%d\n" % k + q + b"\n" for k, (r, q) in enumerate(pp)))
    p = api.parse_fastx(str(fa))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    for algo in (api.LNW, api.ANW, api.LSW):
        res = eng.align_batch(api.make_params(algo, flags=ALL), p.sequences, p.pairs)
        s, e, t = ol.align_batch(ol.params(algo), blob, pairs)
        assert (res.scores == s).all() and res.strings == t


@pytest.mark.parametrize("L", [84, 85, 169, 170, 340, 341])
def test_short_read_keys_at_the_position_bit_thresholds(eng, L):
    """The end-cell keys are unsigned 16-bit: (3 L + 2) << k < 65536 picks k = 8 / 7 / 6 / 5 at L <= 84 / 169 / 340 / ...  Identical
    and nearly identical pairs drive the score -- and with it bit 15 of the keys -- to the top of each range, on both sides of a
    threshold; shifted copies put the maximum into every lane and block."""
    rng = np.random.default_rng(L)
    seqs = []
    for k in range(96):
        r = rng.integers(0, 4, L).astype(np.uint8) + ord("0")
        q = r.copy()
        if k % 3 == 1:
            q[rng.integers(0, L, max(1, L // 40))] = ord("0")
        if k % 3 == 2:
            q = np.concatenate([q[k % 17:], rng.integers(0, 4, k % 17).astype(np.uint8) + ord("0")])
        seqs.append((r.tobytes(), q.tobytes()))
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(seqs))
    w = dict(match=3, mismatch=-1, gap_open=-2)
    res, st = _staged(eng, blob, pairs, api.make_params(api.LSW, flags=api.OUT_SCORE | api.OUT_END_COORDS, **w))
    assert st["kernel_id"] == 2
    s, e, _ = ol.align_batch(ol.params(ol.LSW, **w), blob, pairs, strings=False, threads=8)
    assert (res.scores == s).all() and int(s.max()) == 3 * L
    assert (res.end_row_col == e).all()
