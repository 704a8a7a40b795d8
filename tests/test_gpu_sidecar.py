"""The packed-sidecar upload (csrc/host_pack.cpp + batch_from_sidecar) and the multi-device call (csrc/host_multi.cpp) against the
oracle, and against the raw-byte upload of the same input (option no_sidecar).  Also the round-1 advisor findings that needed a
GPU to reproduce: concurrent fill chunks with global boundary rows, mixed-case stripes."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import api, longpair, synth

pytestmark = pytest.mark.gpu
ALL = api.OUT_SCORE | api.OUT_END_COORDS | api.OUT_STRINGS
SE = api.OUT_SCORE | api.OUT_END_COORDS
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(api.LNW, dict(gap_open=-2)), (api.ANW, dict(gap_open=-3, gap_extend=-1)), (api.LSW, dict(gap_open=-2)), (api.BSW, dict(gap_open=-2, band=12)),
         (api.LNW, dict(match=3, mismatch=-9, gap_open=-2))]          # last one: weights outside the score tables -> byte-compare kernel


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _ragged_image(n, lo, hi, seed, alphabet=b"0123"):
    rng = synth.Rng(seed)
    pp = [(b"", b""), (alphabet[:1], b""), (b"", alphabet[:2])]
    lens = lo + rng.below(n, hi - lo + 1)
    for k in range(n):
        r = synth.random_seq(rng, int(lens[k]), alphabet)
        pp.append((r, synth.mutate(rng, r, 0.05, 0.02, 0.02, alphabet)))
    return synth.pairs_to_file_bytes(pp)


def _same(res, s, e, t, algo):
    assert (res.scores == s).all()
    if algo in (api.LSW, api.BSW):
        assert (res.end_row_col == e).all()
    if t is not None:
        assert res.strings == t


@pytest.mark.parametrize("alphabet", [b"0123", b"ACGT", b"ca"])
@pytest.mark.parametrize("algo,w", CASES)
def test_sidecar_upload_matches_oracle_and_raw_upload(eng, algo, w, alphabet):
    inp = api.parse_image_native(_ragged_image(120, 1, 330, 7, alphabet))
    assert api.input_sidecar(inp.sequences) is not None
    s, e, t = ol.align_batch(ol.params(algo, **w), inp.sequences, inp.pairs)
    _same(eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs), s, e, t, algo)
    _same(eng.align_batch(api.make_params(algo, flags=SE, **w), inp.sequences, inp.pairs), s, e, None, algo)
    with eng.options(no_sidecar=1):
        _same(eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs), s, e, t, algo)
    # any sub-range of the registered index is served from the same sidecar
    sub = slice(17, 90)
    res = eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs[sub])
    _same(res, s[sub], e[sub], t[sub], algo)
    assert eng.align_batch_text(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs, 5) == ol.format_text(s, t, 5)
    inp.free()


def test_golden_text_through_the_native_parser(eng):
    for name, algo, tag, w in (("cfg1_small", api.LNW, "LNW", dict(gap_open=-2)), ("adversarial", api.ANW, "ANW", dict(gap_open=-3, gap_extend=-1)),
                               ("mid", api.LSW, "LSW", dict(gap_open=-2))):
        inp = api.parse_input_native(os.path.join(GOLD, f"{name}.in.txt"))
        got = eng.align_batch_text(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs)
        assert got == open(os.path.join(GOLD, f"{name}.{tag}.out.txt"), "rb").read()
        inp.free()


@pytest.mark.parametrize("kind", ["uniform", "ragged"])
def test_chunked_one_call_route_from_the_sidecar(eng, kind):
    """>= 65 536 pairs: the score / end-cell request is cut into chunks over four streams, every chunk uploading its slice of the
    sidecar (uniform: packed words only; ragged: + sizes and word offsets)."""
    n = 140_000
    if kind == "uniform":
        inp = api.parse_image_native(synth.uniform_file_bytes(n, 100, 90, 21))
    else:
        blob, pairs = synth.ragged_mutated_blob_pairs(n, 40, 160, 22, 0.05, 0.02, 0.02)
        inp = api.parse_image_native(synth.blob_to_file_bytes(blob))
    sc = api.input_sidecar(inp.sequences)
    assert sc["uniform"] == (kind == "uniform")
    s, e, _ = ol.align_batch(ol.params(ol.LSW), inp.sequences, inp.pairs, strings=False, threads=8)
    res = eng.align_batch(api.make_params(api.LSW, flags=SE), inp.sequences, inp.pairs)
    _same(res, s, e, None, api.LSW)
    with eng.options(no_sidecar=1):
        _same(eng.align_batch(api.make_params(api.LSW, flags=SE), inp.sequences, inp.pairs), s, e, None, api.LSW)
    lo = 33_333
    _same(eng.align_batch(api.make_params(api.LSW, flags=SE), inp.sequences, inp.pairs[lo:]), s[lo:], e[lo:], None, api.LSW)
    inp.free()


@pytest.mark.parametrize("algo,w", [(api.ANW, dict(gap_open=-3, gap_extend=-1)), (api.LNW, dict(gap_open=-2)), (api.LSW, dict(gap_open=-2))])
def test_concurrent_fill_chunks_with_global_boundary_rows(eng, algo, w):
    """references long enough for the per-warp GLOBAL boundary rows (R ~ 2000), queries of several passes
    (Q > 256) and a traceback budget that forces >= 2 chunks over the two slab buffers: fills c and c+1 run concurrently and
    must not share boundary rows."""
    img = synth.mutated_fixed_file_bytes(48, 2000, 620, 31, 0.04, 0.01, 0.01)
    blob, pairs = ol.parse_image(img)
    s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs, threads=8)
    eng.set_option("tb_budget_bytes", 8 << 20)
    try:
        for _ in range(3):
            b = eng.upload(blob, pairs)
            b.run(api.make_params(algo, flags=ALL, **w))
            st_launches = None
            res = b.fetch()
            st_launches = b.stats()["kernel_launches"]
            b.free()
            assert st_launches >= 4, "the budget did not force several chunks"
            _same(res, s, e, t, algo)
    finally:
        eng.set_option("tb_budget_bytes", 16 << 30)


def test_stripe_mode_compares_bytes_exactly_like_the_reference(eng):
    """a soft-masked (lower-case) reference against upper-case reads must NOT match in stripe mode either."""
    rng = synth.Rng(3)
    up = synth.random_seq(rng, 6000, b"ACGT")
    qry = synth.mutate(rng, up, 0.02, 0.005, 0.005, b"ACGT")
    mixed = up[:2000] + up[2000:4000].lower() + up[4000:]
    p = api.make_params(api.LSW)
    for ref in (up, up.lower(), mixed):
        want = ol.lsw_score_only(ol.params(ol.LSW), ref, qry)
        assert eng.align_long_pair(p, ref, qry) == want
        job = longpair.StripedLongPair(eng, p, ref, qry, 0, 1, None)
        got, _ = job.run()
        job.free()
        assert got == want, (got, want)


# ---- several devices behind one host process (dpx_create_multi) ----------------------------------------------------------
def _device_sets():
    import torch
    n = torch.cuda.device_count()
    # several workers on one GPU: the sharding / stitching code without needing several GPUs (eight: the few-large-chunks setting)
    sets = [[0], [0, 0, 0], [0] * 8]
    if n > 1:
        sets.append(list(range(n)))
    return sets


@pytest.mark.parametrize("devices", _device_sets())
def test_multi_device_call_returns_pair_order_results(eng, devices):
    m = api.MultiEngine(devices=devices)
    assert m.n_devices == len(devices)
    inp = api.parse_image_native(_ragged_image(700, 1, 300, 41))
    for algo, w in CASES[:4]:
        s, e, t = ol.align_batch(ol.params(algo, **w), inp.sequences, inp.pairs, threads=8)
        _same(m.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs), s, e, t, algo)
        _same(m.align_batch(api.make_params(algo, flags=SE, **w), inp.sequences, inp.pairs), s, e, None, algo)
        assert m.align_batch_text(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs, 3) == ol.format_text(s, t, 3)
    # fewer pairs than devices: empty shards
    one = eng.align_batch(api.make_params(api.LNW, flags=ALL), inp.sequences, inp.pairs[5:6])
    got = m.align_batch(api.make_params(api.LNW, flags=ALL), inp.sequences, inp.pairs[5:6])
    assert (got.scores == one.scores).all() and got.strings == one.strings
    # unregistered numpy input (raw-byte upload on every device) and the chunked score route
    blob, pairs = synth.uniform_blob_pairs(150_000, 80, 80, 9)
    s, e, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False, threads=8)
    _same(m.align_batch(api.make_params(api.LSW, flags=SE), blob, pairs), s, e, None, api.LSW)
    inp.free()
    m.close()


@pytest.mark.parametrize("algo,w", CASES[:4])
def test_pipelined_strings_route_matches_the_serial_one(eng, algo, w):
    """>= 16 384 pairs with alignment strings: the one-call ABI cuts the batch into chunks over four streams and downloads each
    chunk's strings while the next chunks compute; blob + offsets must spell the same strings as the serial route and the oracle."""
    n = 20000
    blob, pairs = synth.ragged_mutated_blob_pairs(n, 30, 140, 77 + algo, 0.05, 0.02, 0.02)
    inp = api.parse_image_native(synth.blob_to_file_bytes(blob))
    s, e, t = ol.align_batch(ol.params(algo, **w), inp.sequences, inp.pairs, threads=8)
    res = eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs)          # sidecar chunks
    _same(res, s, e, t, algo)
    res = eng.align_batch(api.make_params(algo, flags=ALL, **w), blob, pairs)                       # raw-byte chunks
    _same(res, s, e, t, algo)
    with eng.options(serial_strings=1):
        _same(eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs), s, e, t, algo)
    inp.free()
    # uniform lengths (no schedule), sub-range of a registered input
    img = synth.mutated_fixed_file_bytes(18000, 120, 110, 5, 0.04, 0.01, 0.01)
    inp = api.parse_image_native(img)
    s, e, t = ol.align_batch(ol.params(algo, **w), inp.sequences, inp.pairs, threads=8)
    _same(eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs), s, e, t, algo)
    _same(eng.align_batch(api.make_params(algo, flags=ALL, **w), inp.sequences, inp.pairs[700:17700]), s[700:17700], e[700:17700], t[700:17700], algo)
    inp.free()


def test_stale_registration_is_not_used(eng):
    """The packed copy belongs to the bytes that were registered: if the blob's head / tail or the sizes of the first / last pair of
    the call no longer match (memory reused for another input at the same address, or edited in place), uploads fall back to the bytes."""
    blob, pairs = synth.uniform_blob_pairs(300, 60, 60, 77)
    blob = blob.copy(); pairs = pairs.copy()
    api.register_input(blob, pairs)
    p = api.make_params(api.LSW, flags=SE)
    s0, e0, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False)
    _same(eng.align_batch(p, blob, pairs), s0, e0, None, api.LSW)
    blob[pairs["referenceIdx"][0]: pairs["referenceIdx"][0] + 60] = blob[pairs["queryIdx"][0]: pairs["queryIdx"][0] + 60]    # pair 0 now matches itself
    s1, e1, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False)
    assert s1[0] == 180 and s1[0] != s0[0]
    _same(eng.align_batch(p, blob, pairs), s1, e1, None, api.LSW)
    pairs["querySize"][-1] = 31                                                       # an edited index (last pair of the call)
    s2, e2, _ = ol.align_batch(ol.params(ol.LSW), blob, pairs, strings=False)
    _same(eng.align_batch(p, blob, pairs), s2, e2, None, api.LSW)
    api.unregister_input(blob)
