"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/dpx_oracle.c).

Also holds a numpy restatement of the reference parser (c++/parseInput.cpp:78-113) used by
the tests to turn a file image into (blob, seqPair[]) without touching the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_ALIGN = os.path.join(ORACLE_DIR, "_ref", "ref_align")

LNW, ANW, LSW, BSW, ABSW = 0, 1, 2, 3, 4
ALGO_NAMES = {LNW: "LNW", ANW: "ANW", LSW: "LSW", BSW: "BSW", ABSW: "ABSW"}


class OrcParams(C.Structure):
    _fields_ = [("algo", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32),
                ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("band", C.c_int32)]


PAIR_DTYPE = np.dtype([("referenceIdx", "<i4"), ("referenceSize", "<i4"), ("queryIdx", "<i4"), ("querySize", "<i4")])

_lib = None


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ORACLE_DIR, "dpx_oracle.c")):
            build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_align_batch.restype = C.c_int
        _lib.orc_align_batch.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_lsw_all_text.restype = C.c_int
        _lib.orc_lsw_all_text.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
        _lib.orc_free.restype = None; _lib.orc_free.argtypes = [C.c_void_p]
        _lib.orc_lsw_score_only.restype = C.c_int
        _lib.orc_lsw_score_only.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    return _lib


def parse_image(img: bytes | np.ndarray):
    """(blob uint8[numBytes] with '\\n'->0, pairs PAIR_DTYPE[numPairs]) — c++/parseInput.cpp:78-113."""
    a = np.frombuffer(img, dtype=np.uint8) if isinstance(img, (bytes, bytearray)) else np.asarray(img, dtype=np.uint8)
    nl = np.flatnonzero(a == 10)
    if len(nl) % 3 != 0:
        raise ValueError("Number of lines not a multiple of 3")  # parseInput.cpp:38-41
    n = len(nl) // 3
    blob = a.copy()
    blob[nl] = 0
    pairs = np.zeros(n, dtype=PAIR_DTYPE)
    h, r, q = nl[0::3], nl[1::3], nl[2::3]
    pairs["referenceIdx"] = h + 1
    pairs["referenceSize"] = r - (h + 1)
    pairs["queryIdx"] = r + 1
    pairs["querySize"] = q - (r + 1)
    return blob, pairs


def params(algo, match=3, mismatch=-1, gap_open=-2, gap_extend=-1, band=0) -> OrcParams:
    return OrcParams(algo, match, mismatch, gap_open, gap_extend, band)


def align_batch(p: OrcParams, blob: np.ndarray, pairs: np.ndarray, strings: bool = True, threads: int = 1,
                bandmem: bool = False):
    """Returns (scores int32[n], end_rc int32[n,2], list of (REF, REL, QRY) bytes or None)."""
    n = len(pairs)
    scores = np.zeros(n, dtype=np.int32)
    end_rc = np.zeros((n, 2), dtype=np.int32)
    blob = np.ascontiguousarray(blob)
    pairs = np.ascontiguousarray(pairs)
    sbuf = None
    off = None
    if strings:
        f = pairs["referenceSize"].astype(np.int64) + pairs["querySize"].astype(np.int64) + 1
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(3 * f, out=off[1:])
        sbuf = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
    lib().orc_align_batch(C.byref(p), blob.ctypes.data, pairs.ctypes.data, n, threads,
                          scores.ctypes.data, end_rc.ctypes.data,
                          sbuf.ctypes.data if strings else None, off.ctypes.data if strings else None, int(bandmem))
    out = None
    if strings:
        out = []
        raw = sbuf.tobytes()
        for i in range(n):
            fi = int(f[i]); o = int(off[i])
            trip = []
            for k in range(3):
                s = raw[o + k * fi: o + (k + 1) * fi]
                trip.append(s[: s.index(b"\0")])
            out.append(tuple(trip))
    return scores, end_rc, out


def format_text(scores, strs, first_index: int = 0) -> bytes:
    """Reference stdout blocks: '<i> | <score>\\nREF\\nREL\\nQRY\\n' (c++/LinearNeedlemanWunsch.cpp:207-213)."""
    parts = []
    for i, (s, (a, b, c)) in enumerate(zip(scores, strs)):
        parts.append(b"%d | %d\n" % (first_index + i, int(s)))
        parts.append(a + b"\n" + b + b"\n" + c + b"\n")
    return b"".join(parts)


def lsw_score_only(p: OrcParams, ref: bytes, qry: bytes, band: int = -1):
    s = C.c_int32(); r = C.c_int64(); c = C.c_int64()
    lib().orc_lsw_score_only(C.byref(p), band, ref, len(ref), qry, len(qry), C.byref(s), C.byref(r), C.byref(c))
    return s.value, r.value, c.value


def lsw_all_text(p: OrcParams, blob: np.ndarray, pairs: np.ndarray, first_index: int = 0):
    """LinearSmithWaterman with BACKTRACK_ALL: (stdout blocks with one alignment per maximum cell, number of alignments)."""
    blob = np.ascontiguousarray(blob); pairs = np.ascontiguousarray(pairs)
    txt = C.c_void_p(); nb = C.c_size_t(); na = C.c_int64()
    lib().orc_lsw_all_text(C.byref(p), blob.ctypes.data, pairs.ctypes.data, len(pairs), first_index, C.byref(txt), C.byref(nb), C.byref(na))
    try:
        return C.string_at(txt, nb.value), na.value
    finally:
        lib().orc_free(txt)


REF_ALIGN_ALL = os.path.join(os.path.dirname(REF_ALIGN), "ref_align_all")


def run_reference_all(path: str, match=3, mismatch=-1, gap_open=-2) -> bytes:
    """stdout blocks of the reference's LinearSmithWaterman compiled with -DBACKTRACK_ALL (oracle/_ref/ref_align_all)."""
    cmd = [REF_ALIGN_ALL, "-algo", "LSW", "-pairs", path, "-match", str(match), "-mismatch", str(mismatch), "-open", str(gap_open), "-noheader"]
    return subprocess.run(cmd, check=True, capture_output=True).stdout


def have_ref_binary() -> bool:
    return os.path.exists(REF_ALIGN)


def run_reference(algo: int, path: str, match=3, mismatch=-1, gap_open=-2, gap_extend=-1, threads=1) -> bytes:
    """stdout blocks of the compiled unmodified reference classes (oracle/_ref/ref_align -noheader)."""
    cmd = [REF_ALIGN, "-algo", ALGO_NAMES[algo], "-pairs", path, "-match", str(match), "-mismatch", str(mismatch),
           "-open", str(gap_open), "-extend", str(gap_extend), "-threads", str(threads), "-noheader"]
    return subprocess.run(cmd, check=True, capture_output=True).stdout
