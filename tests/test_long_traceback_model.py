"""CPU model of the long-pair traceback scheme (csrc/longtrace.cuh), checked against the oracle's full-matrix backtrack.

The GPU keeps no traceback for a 1 Mbp x 1 Mbp matrix.  It keeps CHECKPOINTS: H on every TW-th column (what each warp of the
forward kernel reads from its left neighbour) and on every TH-th row (what each lane holds when it crosses such a row).  The walk
then re-fills one TH x TW tile at a time from its top row and left column, with the reference's direction rule
(c++/LinearSmithWaterman.cpp:104-108), follows the directions to the tile's edge and moves on.  This file restates that scheme in
numpy at toy tile sizes: it pins the checkpoint indexing, the tile-edge handling and the per-tile transfer tables (exit cell and
move count of a walk entering through the tile's last row or column) that let the GPU hop over tiles without reading directions.
The CUDA kernels are a transcription of `walk` / `transfer` below and are tested on the GPU against the same oracle
(tests/test_gpu_longtrace.py).  Test infrastructure only — nothing here is on the product path."""
import numpy as np
import pytest

import oracle_lib as ol
from dpx_gpu_genomics_project_b200 import synth

LSW = 2
STOP, UP, LEFT, DIAG = 0, 1, 2, 3


def full_h(ref, qry, m, x, g):
    Q, R = len(qry), len(ref)
    H = np.zeros((Q + 1, R + 1), dtype=np.int64)
    for i in range(1, Q + 1):
        for j in range(1, R + 1):
            s = m if qry[i - 1] == ref[j - 1] else x
            H[i, j] = max(0, H[i - 1, j - 1] + s, H[i - 1, j] + g, H[i, j - 1] + g)
    return H


def walk(ref, qry, m, x, g, ie, je, colck, rowck, TH, TW):
    """colck[c][i] = H[i][(c+1)*TW], rowck[r][j] = H[(r+1)*TH][j].  Returns (REF, REL, QRY, start_row, start_col)."""
    out = [bytearray(), bytearray(), bytearray()]
    i, j = ie, je
    tiles = 0
    while True:
        tr, tc = (i - 1) // TH, (j - 1) // TW
        r0, c0 = tr * TH, tc * TW
        h, w = i - r0, j - c0
        # borders of the tile: row r0 (w+1 values from column c0) and column c0 (h+1 values from row r0)
        top = np.zeros(w + 1, dtype=np.int64) if tr == 0 else rowck[tr - 1][c0:c0 + w + 1].copy()
        left = np.zeros(h + 1, dtype=np.int64) if tc == 0 else colck[tc - 1][r0:r0 + h + 1].copy()
        assert top[0] == left[0]
        T = np.zeros((h + 1, w + 1), dtype=np.int64)
        D = np.zeros((h + 1, w + 1), dtype=np.int8)
        T[0, :] = top; T[:, 0] = left
        for a in range(1, h + 1):
            for b in range(1, w + 1):
                s = m if qry[r0 + a - 1] == ref[c0 + b - 1] else x
                up, lf, dg = T[a - 1, b] + g, T[a, b - 1] + g, T[a - 1, b - 1] + s
                v = max(0, up, lf, dg)
                T[a, b] = v
                D[a, b] = STOP if v == 0 else UP if up == v else LEFT if lf == v else DIAG
        tiles += 1
        a, b = h, w
        stopped = False
        while a > 0 and b > 0:
            d = D[a, b]
            if d == STOP:
                stopped = True
                break
            qi, rj = qry[r0 + a - 1], ref[c0 + b - 1]
            if d == DIAG:
                out[0].append(rj); out[1].append(ord("*") if qi == rj else ord("|")); out[2].append(qi); a -= 1; b -= 1
            elif d == UP:
                out[0].append(ord("_")); out[1].append(ord(" ")); out[2].append(qi); a -= 1
            else:
                out[0].append(rj); out[1].append(ord(" ")); out[2].append(ord("_")); b -= 1
        i, j = r0 + a, c0 + b
        if stopped or i == 0 or j == 0:
            break
        # on a tile edge: the cell (i, j) itself belongs to the next tile; its H decides whether the walk goes on
        # (c++/LinearSmithWaterman.cpp:222 stops when H[next] == 0) — that tile's fill answers it through its STOP code
    return bytes(out[0][::-1]), bytes(out[1][::-1]), bytes(out[2][::-1]), i, j, tiles


@pytest.mark.parametrize("seed,R,Q,TH,TW,w", [
    (1, 150, 170, 16, 24, (3, -1, -2)), (2, 97, 64, 8, 8, (1, -1, -1)), (3, 200, 40, 32, 16, (2, -3, -2)),
    (4, 64, 64, 16, 16, (3, -1, -2)), (5, 130, 131, 128, 128, (3, -1, -2)), (6, 90, 140, 7, 5, (5, -4, -3)),
])
def test_checkpointed_tile_walk_equals_full_matrix_backtrack(seed, R, Q, TH, TW, w):
    m, x, g = w
    rng = synth.Rng(seed)
    ref = synth.random_seq(rng, R, b"012")
    qry = synth.mutate(rng, ref, 0.08, 0.04, 0.04, b"012")[:Q]
    qry = qry + synth.random_seq(rng, Q - len(qry), b"012")
    blob, pairs = ol.parse_image(synth.pairs_to_file_bytes([(ref, qry)]))
    s, e, t = ol.align_batch(ol.params(LSW, match=m, mismatch=x, gap_open=g), blob, pairs)
    H = full_h(ref, qry, m, x, g)
    assert H.max() == s[0]
    ie, je = int(e[0][0]), int(e[0][1])
    if s[0] == 0:
        return
    colck = [H[:, c] for c in range(TW, R + 1, TW)]
    rowck = [H[r, :] for r in range(TH, Q + 1, TH)]
    a, b, c, i0, j0, tiles = walk(ref, qry, m, x, g, ie, je, colck, rowck, TH, TW)
    assert (a, b, c) == t[0]
    assert tiles >= 1 and H[i0, j0] == 0


def transfer(D):
    """Exit cell (tile-relative, -1 = the row above / column left of the tile) and move count of the walk from every cell of a tile
    whose directions are D[1..h][1..w] — the recurrence the fill kernel carries beside H: a cell takes over its predecessor's answer."""
    h, w = D.shape[0] - 1, D.shape[1] - 1
    E = {}
    for a in range(1, h + 1):
        for b in range(1, w + 1):
            d = D[a, b]
            if d == STOP:
                E[a, b] = (a - 1, b - 1, 0, True)
                continue
            pa, pb = (a - 1, b - 1) if d == DIAG else (a - 1, b) if d == UP else (a, b - 1)
            if pa == 0 or pb == 0:
                E[a, b] = (pa - 1, pb - 1, 1, False)
            else:
                x, y, n, st = E[pa, pb]
                E[a, b] = (x, y, n + 1, st)
    return E


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_transfer_table_equals_walking_the_directions(seed):
    rng = np.random.default_rng(seed)
    h, w = 13, 17
    D = rng.integers(0, 4, size=(h + 1, w + 1)).astype(np.int8)
    D[rng.random(D.shape) < 0.7] = DIAG
    E = transfer(D)
    for a, b in [(h, y) for y in range(1, w + 1)] + [(x, w) for x in range(1, h + 1)]:
        x, y, n, stopped = a, b, 0, False
        while x > 0 and y > 0:
            d = D[x, y]
            if d == STOP:
                stopped = True
                break
            x, y = (x - 1, y - 1) if d == DIAG else (x - 1, y) if d == UP else (x, y - 1)
            n += 1
        assert E[a, b] == (x - 1, y - 1, n, stopped)
