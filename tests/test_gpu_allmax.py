"""LinearSmithWaterman in the reference's BACKTRACK_ALL mode on the GPU (csrc/allmax.cuh, dpx_align_batch_text_all) against the oracle
restatement, which tests/test_oracle.py pins byte-for-byte on the reference compiled with -DBACKTRACK_ALL."""
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


def _tie_heavy(seed, n, alphabet, hi):
    rng = synth.Rng(seed)
    pp = []
    for k in range(n):
        R, Q = 1 + int(rng.below(1, hi)[0]), 1 + int(rng.below(1, hi)[0])
        kind = k % 4
        if kind == 0:
            unit = synth.random_seq(rng, 1 + int(rng.below(1, 4)[0]), alphabet)
            pp.append(((unit * hi)[:R], (unit * hi)[1:1 + Q]))
        elif kind == 1:
            pp.append((synth.random_seq(rng, R, alphabet), synth.random_seq(rng, Q, alphabet)))
        elif kind == 2:
            core = synth.random_seq(rng, 2 + int(rng.below(1, 6)[0]), alphabet)
            pp.append((synth.random_seq(rng, R // 3, alphabet) + core + synth.random_seq(rng, R // 3, alphabet) + core,
                       core + synth.random_seq(rng, Q // 2, alphabet) + core))
        else:
            r = synth.random_seq(rng, R, alphabet)
            pp.append((r, synth.mutate(rng, r, 0.1, 0.05, 0.05, alphabet)))
    return pp


@pytest.mark.parametrize("seed,alphabet,hi,w", [(1, b"01", 40, (3, -1, -2)), (2, b"0", 60, (1, -1, -1)), (3, b"012", 90, (2, -3, -2)),
                                                (4, b"0123", 300, (3, -1, -2)), (5, b"ACGTN", 150, (5, -4, -3))])
def test_every_maximum_is_walked_in_the_reference_order(eng, seed, alphabet, hi, w):
    m, x, g = w
    pp = _tie_heavy(seed, 150, alphabet, hi)
    pp += [(b"", b""), (alphabet[:1] * 5, b""), (b"", alphabet[:1] * 3), (b"0" * 7, b"1" * 9)]          # empty and zero-score pairs
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes(pp))
    want, n_want = ol.lsw_all_text(ol.params(ol.LSW, match=m, mismatch=x, gap_open=g), blob, pairs, 5)
    got, n_got = eng.align_batch_text_all(api.make_params(api.LSW, match=m, mismatch=x, gap_open=g), blob, pairs, 5)
    assert n_got == n_want and n_got > len(pp) - 4
    assert got == want


def test_single_maximum_gives_the_ordinary_output_and_other_algorithms_are_refused(eng):
    rng = synth.Rng(9)
    r = synth.random_seq(rng, 400); q = synth.mutate(rng, r, 0.03, 0.01, 0.01)
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(r, q)]))
    got, n = eng.align_batch_text_all(api.make_params(api.LSW), blob, pairs)
    s, e, t = ol.align_batch(ol.params(ol.LSW), blob, pairs)
    if n == 1:
        assert got == ol.format_text(s, t)
    with pytest.raises(api.DpxError):
        eng.align_batch_text_all(api.make_params(api.LNW), blob, pairs)
    assert eng.align_batch_text_all(api.make_params(api.LSW), np.zeros(0, np.uint8), np.zeros(0, api.PAIR_DTYPE)) == (b"", 0)


def test_golden_output_of_the_reference_built_with_backtrack_all(eng):
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    blob, pairs = ol.parse_image(open(os.path.join(gold, "ties.pairs.txt"), "rb").read())
    got, n = eng.align_batch_text_all(api.make_params(api.LSW), blob, pairs)
    assert got == open(os.path.join(gold, "ties.LSW_ALL.out.txt"), "rb").read()
