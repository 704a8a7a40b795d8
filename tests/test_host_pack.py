"""CPU-side checks of the host code added around the C ABI: the in-memory parser (dpx_parse_image), the packed 2-bit sidecar it
registers (csrc/host_pack.cpp) and the shard boundaries of the multi-GPU call (csrc/host_multi.cpp).  No device needed: without
one the sidecar simply lives in pageable memory."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import _lib, api, synth


@pytest.fixture(scope="module", autouse=True)
def L():
    _lib.build()
    return _lib.lib()


def decode(sc, pairs):
    """(ref, qry) bytes of every pair, rebuilt from the sidecar's packed words."""
    inv = np.frombuffer(sc["code_to_byte"], dtype=np.uint8)
    out = []
    for p, pr in enumerate(pairs):
        w = sc["words"][sc["word_offsets"][p]: sc["word_offsets"][p + 1]]
        R, Q = int(pr["referenceSize"]), int(pr["querySize"])
        rw = (R + 15) // 16
        codes = ((w[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).astype(np.uint8)
        out.append((inv[codes[:rw].reshape(-1)[:R]].tobytes(), inv[codes[rw:].reshape(-1)[:Q]].tobytes()))
    return out


def raw_pairs(blob, pairs):
    b = blob.tobytes()
    return [(b[p["referenceIdx"]: p["referenceIdx"] + p["referenceSize"]], b[p["queryIdx"]: p["queryIdx"] + p["querySize"]]) for p in pairs]


@pytest.mark.parametrize("alphabet", [b"0123", b"ACGT", b"acgt", b"AT", b"7", b"xyzw"])
def test_parse_image_and_sidecar_round_trip(alphabet):
    rng = synth.Rng(11)
    pp = [(b"", b""), (alphabet[:1] * 16, alphabet[-1:] * 17), (alphabet[:1], b"")]
    for n in (1, 15, 16, 17, 31, 33, 150, 1000):
        r = synth.random_seq(rng, n, alphabet)
        pp.append((r, synth.mutate(rng, r, 0.1, 0.05, 0.05, alphabet)))
    img = synth.pairs_to_file_bytes(pp)
    inp = api.parse_image_native(img)
    blob, pairs = ol.parse_image(img)
    assert (inp.sequences == blob).all() and (inp.pairs == pairs).all()
    assert inp.info["numPairs"] == len(pp) and inp.info["numCells"] == sum(len(a) * len(b) for a, b in pp)
    sc = api.input_sidecar(inp.sequences)
    assert sc is not None and sc["n_pairs"] == len(pp) and sc["n_symbols"] == len(set(alphabet)) and not sc["uniform"]
    assert decode(sc, pairs) == raw_pairs(blob, pairs) == pp
    seqs = inp.sequences
    inp.free()                                    # dpx_free of the blob / index drops the sidecar
    assert _lib.lib().dpx_input_sidecar(seqs.ctypes.data if seqs is not None else 0, None, None, None, None, None, None, None) == 0


def test_uniform_lengths_and_upload_accounting():
    img = synth.uniform_file_bytes(3000, 150, 150, 0x5EED0002)
    inp = api.parse_image_native(img)
    sc = api.input_sidecar(inp.sequences)
    assert sc["uniform"] and sc["upload_bytes"] == 4 * 3000 * 20          # 10 + 10 words per pair, nothing else
    assert sc["upload_bytes"] * 3.7 <= inp.sequences.size      # 150 bases take 10 words: 80 B against 304 B per pair
    assert decode(sc, inp.pairs) == raw_pairs(inp.sequences, inp.pairs)
    blob, pairs = synth.ragged_mutated_blob_pairs(500, 100, 300, 5, 0.05, 0.02, 0.02)
    api.register_input(blob, pairs)
    sc = api.input_sidecar(blob)
    assert not sc["uniform"] and sc["upload_bytes"] == 4 * len(sc["words"]) + 8 * 500 + 4
    assert decode(sc, pairs) == raw_pairs(blob, pairs)
    api.unregister_input(blob)
    assert api.input_sidecar(blob) is None


def test_fifth_symbol_stays_on_the_raw_path_and_bad_index_is_refused():
    img = synth.pairs_to_file_bytes([(b"01234", b"43210")])            # the reference's data sets use '0'..'4'
    inp = api.parse_image_native(img)
    assert api.input_sidecar(inp.sequences) is None
    blob = np.frombuffer(b"0123\x000123\x00", dtype=np.uint8).copy()
    bad = np.array([(0, 4, 5, 9)], dtype=api.PAIR_DTYPE)
    with pytest.raises(api.DpxError):
        api.register_input(blob, bad)


def test_parse_image_format_error():
    with pytest.raises(api.DpxError) as e:
        api.parse_image_native(b"0\n0123\n")
    assert e.value.status == -6


@pytest.mark.parametrize("k", [1, 2, 3, 8])
def test_shard_bounds_are_contiguous_and_balanced(k, L):
    import ctypes as C
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 1000):
        pairs = np.zeros(n, dtype=api.PAIR_DTYPE)
        pairs["referenceSize"] = rng.integers(0, 400, n); pairs["querySize"] = rng.integers(0, 400, n)
        out = (C.c_size_t * (k + 1))()
        assert L.dpx_multi_shard_bounds(pairs.ctypes.data, n, k, out) == 0
        b = list(out)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        w = np.maximum(1, pairs["referenceSize"].astype(np.int64) * pairs["querySize"])
        if n >= 1000:
            tot = [int(w[b[g]: b[g + 1]].sum()) for g in range(k)]
            assert max(tot) - min(tot) <= 2 * int(w.max())
