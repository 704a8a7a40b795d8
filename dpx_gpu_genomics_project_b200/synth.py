"""Deterministic synthetic inputs in the reference's parseInput format.

File layout (reference c++/parseInput.cpp:78-113): repeated
``header\\nREF\\nQRY\\n``; the header line is ignored; the file must end with ``\\n`` and
hold a multiple of 3 lines.  Sequences are digit strings over ``'0'..'4'``
(``0->A 1->T 2->C 3->G 4->U``, reference correct-outputs/LNW/web-scraper-LNW.py:5-12).

PRNG: splitmix64, ``seed = 0x5EED0000 + config#``.
"""
from __future__ import annotations

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """n consecutive outputs of splitmix64 seeded with `seed`, starting at stream index `offset`."""
    with np.errstate(over="ignore"):
        k = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


class Rng:
    """Small sequential wrapper over the splitmix64 stream (buffered)."""

    def __init__(self, seed: int):
        self.seed = seed
        self.pos = 0

    def u64(self, n: int) -> np.ndarray:
        out = splitmix64(self.seed, n, self.pos)
        self.pos += n
        return out

    def below(self, n: int, bound: int) -> np.ndarray:
        return (self.u64(n) % np.uint64(bound)).astype(np.int64)

    def uniform(self, n: int) -> np.ndarray:
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def random_seq(rng: Rng, n: int, alphabet: bytes = b"0123") -> bytes:
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.below(n, len(a))].tobytes() if n > 0 else b""


def mutate(rng: Rng, ref: bytes, sub: float, ins: float, dele: float, alphabet: bytes = b"0123") -> bytes:
    """Query derived from `ref`: per-base substitution / insertion-before / deletion."""
    n = len(ref)
    if n == 0:
        return b""
    u = rng.uniform(n)
    newc = np.frombuffer(alphabet, dtype=np.uint8)[rng.below(n, len(alphabet))]
    insc = np.frombuffer(alphabet, dtype=np.uint8)[rng.below(n, len(alphabet))]
    r = np.frombuffer(ref, dtype=np.uint8)
    out = bytearray()
    for k in range(n):
        x = u[k]
        if x < dele:
            continue
        if x < dele + ins:
            out.append(int(insc[k]))
        if x < dele + ins + sub:
            out.append(int(newc[k]))
        else:
            out.append(int(r[k]))
    return bytes(out)


def pairs_to_file_bytes(pairs: list[tuple[bytes, bytes]]) -> bytes:
    """parseInput-format file image: header line = pair index (ignored by the parser)."""
    parts = []
    for k, (r, q) in enumerate(pairs):
        parts.append(b"%d\n" % k)
        parts.append(r + b"\n")
        parts.append(q + b"\n")
    return b"".join(parts)


def uniform_file_bytes(n_pairs: int, R: int, Q: int, seed: int, alphabet: bytes = b"0123") -> np.ndarray:
    """Vectorised generator for fixed-length i.i.d. pairs (configs 2 and 5 of BASELINE.json).

    Every record is ``b"0\\n" + REF + b"\\n" + QRY + b"\\n"`` so the image is a dense
    (n_pairs, 2+R+1+Q+1) uint8 matrix; returned flattened.
    """
    rec = 2 + R + 1 + Q + 1
    a = np.frombuffer(alphabet, dtype=np.uint8)
    img = np.empty((n_pairs, rec), dtype=np.uint8)
    img[:, 0] = ord("0")
    img[:, 1] = 10
    img[:, 2 + R] = 10
    img[:, rec - 1] = 10
    nb = n_pairs * (R + Q)
    # 16 two-bit draws per 64-bit output when the alphabet has 4 letters; else one draw per output
    if len(a) == 4:
        words = splitmix64(seed, (nb + 31) // 32)
        sh = np.arange(0, 64, 2, dtype=np.uint64)
        codes = ((words[:, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[:nb]
    else:
        codes = (splitmix64(seed, nb) % np.uint64(len(a))).astype(np.uint8)
    letters = a[codes].reshape(n_pairs, R + Q)
    img[:, 2:2 + R] = letters[:, :R]
    img[:, 3 + R:3 + R + Q] = letters[:, R:]
    return img.reshape(-1)


def mutated_fixed_file_bytes(n_pairs: int, R: int, Q: int, seed: int, sub: float, ins: float, dele: float,
                             alphabet: bytes = b"0123") -> np.ndarray:
    """Fixed-length pairs whose query is the reference mutated (sub/ins/del) and then trimmed or
    padded with random bases to exactly Q (configs 3 and 4 of BASELINE.json).  Vectorised:
    indels are realised by a per-base source-index walk (cumulative shift)."""
    rec = 2 + R + 1 + Q + 1
    a = np.frombuffer(alphabet, dtype=np.uint8)
    rng = Rng(seed)
    img = np.empty((n_pairs, rec), dtype=np.uint8)
    img[:, 0] = ord("0"); img[:, 1] = 10; img[:, 2 + R] = 10; img[:, rec - 1] = 10
    chunk = max(1, min(n_pairs, (1 << 24) // max(R, Q)))
    for s in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - s)
        ref = a[rng.below(n * R, len(a))].reshape(n, R)
        u = rng.uniform(n * Q).reshape(n, Q)
        # step of the source index per output base: 0 = insertion, 1 = copy, 2 = skip one (deletion)
        step = np.ones((n, Q), dtype=np.int64)
        step[u < ins] = 0
        step[(u >= ins) & (u < ins + dele)] = 2
        src = np.cumsum(step, axis=1) - 1
        rnd = a[rng.below(n * Q, len(a))].reshape(n, Q)
        qry = np.where((src >= 0) & (src < R) & (step > 0), np.take_along_axis(ref, np.clip(src, 0, R - 1), axis=1), rnd)
        subm = rng.uniform(n * Q).reshape(n, Q) < sub
        qry = np.where(subm, rnd, qry)
        img[s:s + n, 2:2 + R] = ref
        img[s:s + n, 3 + R:3 + R + Q] = qry
    return img.reshape(-1)


def uniform_blob_pairs(n_pairs: int, R: int, Q: int, seed: int, alphabet: bytes = b"0123"):
    """Same records as uniform_file_bytes, already in parseInput's output form: (blob with '\\n' -> 0,
    seqPair index) — what parseInput would return for that file (c++/parseInput.cpp:78-113)."""
    img = uniform_file_bytes(n_pairs, R, Q, seed, alphabet)
    rec = 2 + R + 1 + Q + 1
    m = img.reshape(n_pairs, rec)
    m[:, 1] = 0; m[:, 2 + R] = 0; m[:, rec - 1] = 0
    pairs = np.zeros(n_pairs, dtype=[("referenceIdx", "<i4"), ("referenceSize", "<i4"), ("queryIdx", "<i4"), ("querySize", "<i4")])
    base = np.arange(n_pairs, dtype=np.int64) * rec
    pairs["referenceIdx"] = base + 2
    pairs["referenceSize"] = R
    pairs["queryIdx"] = base + 3 + R
    pairs["querySize"] = Q
    return img, pairs


def mutated_blob_pairs(n_pairs: int, R: int, Q: int, seed: int, sub: float, ins: float, dele: float, alphabet: bytes = b"0123"):
    """mutated_fixed_file_bytes in parseInput's output form: (blob with '\\n' -> 0, seqPair index)."""
    img = mutated_fixed_file_bytes(n_pairs, R, Q, seed, sub, ins, dele, alphabet)
    rec = 2 + R + 1 + Q + 1
    m = img.reshape(n_pairs, rec)
    m[:, 1] = 0; m[:, 2 + R] = 0; m[:, rec - 1] = 0
    pairs = np.zeros(n_pairs, dtype=[("referenceIdx", "<i4"), ("referenceSize", "<i4"), ("queryIdx", "<i4"), ("querySize", "<i4")])
    base = np.arange(n_pairs, dtype=np.int64) * rec
    pairs["referenceIdx"] = base + 2
    pairs["referenceSize"] = R
    pairs["queryIdx"] = base + 3 + R
    pairs["querySize"] = Q
    return img, pairs


def ragged_mutated_blob_pairs(n_pairs: int, r_lo: int, r_hi: int, seed: int, sub: float, ins: float, dele: float, alphabet: bytes = b"0123"):
    """BASELINE config 1 substitute: references of length U[r_lo, r_hi], queries = the reference mutated
    (per-base substitution / insertion / deletion), in parseInput's output form.  Vectorised over the whole batch: every pair is
    generated at the maximum length and cut to its own."""
    a = np.frombuffer(alphabet, dtype=np.uint8)
    rng = Rng(seed)
    R = r_lo + rng.below(n_pairs, r_hi - r_lo + 1)
    Rm = int(R.max())
    Qm = Rm + Rm // 4 + 8
    ref = a[rng.below(n_pairs * Rm, len(a))].reshape(n_pairs, Rm)
    u = rng.uniform(n_pairs * Qm).reshape(n_pairs, Qm)
    step = np.ones((n_pairs, Qm), dtype=np.int64)
    step[u < ins] = 0
    step[(u >= ins) & (u < ins + dele)] = 2
    src = np.cumsum(step, axis=1) - 1                       # reference index each query base copies (insertions repeat it)
    rnd = a[rng.below(n_pairs * Qm, len(a))].reshape(n_pairs, Qm)
    subm = rng.uniform(n_pairs * Qm).reshape(n_pairs, Qm) < sub
    qry = np.where((step > 0) & ~subm, np.take_along_axis(ref, np.clip(src, 0, Rm - 1), axis=1), rnd)
    Q = np.maximum(1, (src < R[:, None]).sum(axis=1))       # the query ends where the walk leaves the reference
    rec = 2 + R + 1 + Q + 1
    off = np.zeros(n_pairs + 1, dtype=np.int64); np.cumsum(rec, out=off[1:])
    blob = np.zeros(int(off[-1]), dtype=np.uint8)
    blob[off[:-1]] = ord("0")
    pairs = np.zeros(n_pairs, dtype=[("referenceIdx", "<i4"), ("referenceSize", "<i4"), ("queryIdx", "<i4"), ("querySize", "<i4")])
    pairs["referenceIdx"] = off[:-1] + 2; pairs["referenceSize"] = R
    pairs["queryIdx"] = off[:-1] + 3 + R; pairs["querySize"] = Q
    col_r = np.arange(Rm)[None, :]; col_q = np.arange(Qm)[None, :]
    mr = col_r < R[:, None]; mq = col_q < Q[:, None]
    blob[(pairs["referenceIdx"][:, None] + col_r)[mr]] = ref[mr]
    blob[(pairs["queryIdx"][:, None] + col_q)[mq]] = qry[mq]
    return blob, pairs


def blob_to_file_bytes(blob: np.ndarray) -> np.ndarray:
    """Inverse of the parser's newline->NUL rewrite for generator-made blobs (no NUL inside sequences)."""
    out = blob.copy()
    out[out == 0] = 10
    return out
