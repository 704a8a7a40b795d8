"""Randomised differential test of the whole library against the CPU oracle (test infrastructure): random alphabets, lengths,
weights, algorithms, bands and output selections; every result (scores, end cells, strings, formatted text) must be bit-exact.
usage: python tools/fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as ol  # noqa: E402
from dpx_gpu_genomics_project_b200 import api, synth  # noqa: E402

ALPHABETS = [b"0", b"01", b"012", b"0123", b"ACGT", b"01234", b"ACGTN", b"ACGTNacg", b"012345678"]


def make_batch(rng, seed):
    r = synth.Rng(seed)
    alpha = ALPHABETS[int(rng.integers(len(ALPHABETS)))]
    shape = int(rng.integers(6))
    n = int(rng.integers(1, 60))
    lo, hi = [(0, 40), (1, 150), (100, 320), (200, 700), (900, 1400), (2500, 3300)][shape]
    if shape >= 4:
        n = int(rng.integers(1, 6))
    pp = []
    for k in range(n):
        R = int(rng.integers(lo, hi + 1))
        ref = synth.random_seq(r, R, alpha)
        mode = int(rng.integers(4))
        if mode == 0:
            q = synth.random_seq(r, int(rng.integers(lo, hi + 1)), alpha)
        elif mode == 1:
            q = synth.mutate(r, ref, 0.05, 0.02, 0.02, alpha)
        elif mode == 2:
            q = synth.mutate(r, ref, 0.2, 0.1, 0.1, alpha)
        else:
            a = int(rng.integers(0, max(1, R // 2))); q = ref[a:a + int(rng.integers(0, R + 1))]
        pp.append((ref, q))
    return ol.parse_image(synth.pairs_to_file_bytes(pp)), alpha, (lo, hi), n


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    eng = api.Engine(0)
    t_end = time.time() + seconds
    cases = 0
    kernels = {}
    failures = []
    while time.time() < t_end:
        if rng.integers(120) == 0:
            # a big batch: the chunked one-call pipeline (>= 64k pairs, score / end cell) or the two-slab traceback pipeline (>= 32k pairs, strings)
            n = int(rng.integers(33_000, 140_000))
            seed_b = int(rng.integers(1 << 30))
            if rng.integers(2):
                blob, pairs = synth.ragged_mutated_blob_pairs(n, 8, int(rng.integers(20, 70)), seed_b, 0.05, 0.02, 0.02)
            else:
                L = int(rng.integers(10, 60)); blob, pairs = synth.uniform_blob_pairs(n, L, int(rng.integers(10, 60)), seed_b)
            algo = int(rng.integers(4))
            w = dict(match=3, mismatch=-1, gap_open=-2)
            if rng.integers(2):                                # registered input: every chunk uploads its slice of the sidecar
                api.register_input(blob, pairs)
            if algo == api.ANW:
                w.update(gap_open=-3, gap_extend=-1)
            if algo == api.BSW:
                w["band"] = int(rng.choice([3, 16, 40]))
            strings = bool(rng.integers(2))
            flags = api.OUT_SCORE | api.OUT_END_COORDS | (api.OUT_STRINGS if strings else 0)
            tag = f"case {cases} seed {seed} BIG n {n} algo {algo} w {w} strings {strings}"
            try:
                res = eng.align_batch(api.make_params(algo, flags=flags, **w), blob, pairs)
                s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs, strings=strings, threads=16)
                assert (res.scores == s).all(), f"scores, first bad {np.flatnonzero(res.scores != s)[:5]}"
                assert (res.end_row_col == e).all(), "end cells"
                if strings:
                    assert res.strings == t, "strings"
                kernels["big"] = kernels.get("big", 0) + 1
                api.unregister_input(blob)
            except (AssertionError, api.DpxError) as ex:
                api.unregister_input(blob)
                failures.append(f"{type(ex).__name__}: {ex} :: {tag}")
                print("FAIL", failures[-1], flush=True)
            cases += 1
            continue
        (blob, pairs), alpha, span, n = make_batch(rng, int(rng.integers(1 << 30)))
        algo = int(rng.integers(5))
        m = int(rng.integers(1, 7)); x = -int(rng.integers(0, 6)); g = -int(rng.integers(1, 8))
        w = dict(match=m, mismatch=x, gap_open=g)
        if algo == api.ANW:
            w["gap_open"] = -int(rng.integers(0, 8)); w["gap_extend"] = -int(rng.integers(0, 4))
            if w["gap_open"] == 0 and w["gap_extend"] == 0:
                w["gap_extend"] = -1
        if algo == api.BSW:
            w["band"] = int(rng.choice([0, 1, 5, 16, 31, 32, 33, 64, 65, 96, 97, 150]))
        if algo == api.ABSW:                                   # affine banded SW (not a reference algorithm; oracle: absw_pair)
            w["gap_open"] = -int(rng.integers(0, 8)); w["gap_extend"] = -int(rng.integers(0, 4))
            if w["gap_open"] == 0 and w["gap_extend"] == 0:
                w["gap_extend"] = -1
            w["band"] = int(rng.choice([0, 1, 5, 16, 33, 64, 150, 5000]))
        flags = api.OUT_SCORE | api.OUT_END_COORDS | (api.OUT_STRINGS if rng.integers(3) else 0)
        strings = bool(flags & api.OUT_STRINGS)
        tag = f"case {cases} seed {seed} algo {algo} w {w} flags {flags} alpha {alpha} span {span} n {n}"
        try:
            native = None
            if rng.integers(3) == 0:                           # through the library's parser: uploads come from the packed 2-bit sidecar
                native = api.parse_image_native(synth.blob_to_file_bytes(blob) if len(blob) else b"")
                if native.info["numPairs"] == len(pairs):
                    blob, pairs = native.sequences, native.pairs
                    tag += " sidecar" if api.input_sidecar(blob) is not None else " native"
            b = eng.upload(blob, pairs)
            b.run(api.make_params(algo, flags=flags, **w)); b.sync()
            kid = b.stats()["kernel_id"]
            kernels[kid] = kernels.get(kid, 0) + 1
            tag += f" kernel {kid}"
            res = b.fetch()
            b.free()
            s, e, t = ol.align_batch(ol.params(algo, **w), blob, pairs, strings=strings, threads=16)
            assert (res.scores == s).all(), f"scores, first bad {np.flatnonzero(res.scores != s)[:5]}"
            assert (res.end_row_col == e).all(), "end cells"
            if strings:
                assert res.strings == t, "strings"
                if rng.integers(4) == 0:
                    assert eng.align_batch_text(api.make_params(algo, flags=flags, **w), blob, pairs) == ol.format_text(s, t), "text"
            if algo == api.LSW and span[1] <= 320 and rng.integers(4) == 0:
                # the reference's BACKTRACK_ALL mode: every maximum cell walked, blocks in the reference's order
                want_all, n_all = ol.lsw_all_text(ol.params(algo, **w), blob, pairs, 3)
                got_all, n_got = eng.align_batch_text_all(api.make_params(api.LSW, **w), blob, pairs, 3)
                assert n_got == n_all and got_all == want_all, "all-maxima text"
            if algo == api.LSW and len(pairs) and rng.integers(3) == 0:
                # the long-pair entry points on one pair of the batch, at a random lane width / tile geometry / round size
                k = int(rng.integers(len(pairs))); pr = pairs[k]
                ref = blob[int(pr["referenceIdx"]): int(pr["referenceIdx"]) + int(pr["referenceSize"])].tobytes()
                qry = blob[int(pr["queryIdx"]): int(pr["queryIdx"]) + int(pr["querySize"])].tobytes()
                if len(ref) and len(qry):
                    lk, lt, lc = int(rng.choice([0, 2, 4, 8, 16, 32])), int(rng.choice([0, 1, 3, 40])), int(rng.choice([0, 0, 8, 16]))
                    eng.set_option("long_k", lk); eng.set_option("long_bt_tiles", lt); eng.set_option("long_cap", lc)
                    tag += f" long pair {k} options K={lk} tiles={lt} cap={lc}"
                    p1 = api.make_params(api.LSW, **w)
                    s1, e1, t1 = ol.align_batch(ol.params(algo, **w), blob, pairs[k:k + 1], strings=True)
                    want = (int(s1[0]), int(e1[0][0]), int(e1[0][1]))
                    assert eng.align_long_pair(p1, ref, qry) == want, "long pair score / end cell"
                    end, start, lines, st = eng.align_long_pair_strings(p1, ref, qry)
                    assert end == want and lines == t1[0], "long pair strings"
                    for v in ("long_k", "long_bt_tiles", "long_cap"):
                        eng.set_option(v, 0)
        except (AssertionError, api.DpxError) as ex:       # keep going: one run should list every distinct failure
            for v in ("long_k", "long_bt_tiles", "long_cap"):
                eng.set_option(v, 0)
            failures.append(f"{type(ex).__name__}: {ex} :: {tag}")
            print("FAIL", failures[-1], flush=True)
            if len(failures) >= 10:
                break
        cases += 1
    if failures:
        print(f"fuzz FAILED: {len(failures)} of {cases} batches")
        sys.exit(1)
    print(f"fuzz ok: {cases} batches in {seconds:.0f} s, kernel ids used {dict(sorted(kernels.items(), key=str))}")


if __name__ == "__main__":
    main()
