"""Python mirror of the reference's alignment interface on top of the C ABI (include/dpxalign.h).

Same names and argument order as the reference's C++ classes so that parity tests read like the
reference's own usage (c++/main.cpp:237-252):

    LinearNeedlemanWunsch(ref, query, pairNum, match, mismatch, gap).align()      c++/LinearNeedlemanWunsch.h:42-47
    AffineNeedlemanWunsch(ref, query, pairNum, match, mismatch, open, extend)     c++/AffineNeedlemanWunsch.h:59-66
    LinearSmithWaterman(ref, query, pairNum, match, mismatch, gap)                c++/LinearSmithWaterman.h:51-57
    BandedSmithWaterman(ref, query, match, mismatch, gap, pairNum, band_width)    c++/BandedSmithWaterman.h:51-57
                                         (pairNum last as in the reference; band_width is the parameter
                                          the reference forgot to initialise, .h:16)

``align()`` writes exactly the bytes the reference prints (``"<i> | <score>\\nREF\\nREL\\nQRY\\n"``).
Everything computes on the GPU through libdpxalign.so; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass

import numpy as np

from . import _lib

LNW, ANW, LSW, BSW, ABSW = 0, 1, 2, 3, 4
OUT_SCORE, OUT_END_COORDS, OUT_STRINGS = 1, 2, 4
PAIR_DTYPE = np.dtype([("referenceIdx", "<i4"), ("referenceSize", "<i4"), ("queryIdx", "<i4"), ("querySize", "<i4")])


class DpxError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = _lib.lib().dpx_strerror(status).decode()
        super().__init__(f"dpxalign status {status}: {msg}" + (f" ({detail})" if detail else ""))


def _check(st: int, ctx=None):
    if st != 0:
        detail = _lib.lib().dpx_last_error(ctx).decode() if ctx else ""
        raise DpxError(st, detail)


@dataclass
class ParsedInput:
    """parseInput's outputs (c++/parseInput.h:9-35): blob with newlines -> NUL, seqPair index, inputInfo."""
    sequences: np.ndarray     # uint8[numBytes]
    pairs: np.ndarray         # PAIR_DTYPE[numPairs]
    info: dict


def parse_input(path: str) -> ParsedInput:
    L = _lib.lib()
    pairs_p, seq_p, info = C.c_void_p(), C.c_void_p(), _lib.InputInfo()
    _check(L.dpx_parse_input(path.encode(), C.byref(pairs_p), C.byref(seq_p), C.byref(info)))
    try:
        n, nb = info.numPairs, info.numBytes
        seqs = np.ctypeslib.as_array(C.cast(seq_p, C.POINTER(C.c_uint8)), shape=(max(nb, 1),))[:nb].copy()
        pairs = np.frombuffer(C.string_at(pairs_p, n * PAIR_DTYPE.itemsize), dtype=PAIR_DTYPE).copy() if n else np.zeros(0, PAIR_DTYPE)
    finally:
        L.dpx_free(pairs_p); L.dpx_free(seq_p)
    return ParsedInput(seqs, pairs, {f: getattr(info, f) for f, _ in _lib.InputInfo._fields_})


class NativeInput:
    """parseInput's outputs LEFT IN LIBRARY MEMORY (numpy views, no copies): the blob / index pointers are the ones the parser
    registered its packed 2-bit sidecar under, so `Engine.align_batch(params, inp.sequences, inp.pairs)` uploads the packed
    words instead of the bytes (include/dpxalign.h, "Packed sidecar").  Treat both arrays as read-only; call free() when done."""

    def __init__(self, pairs_p, seq_p, info):
        self._L = _lib.lib()
        self._pairs_p, self._seq_p = pairs_p, seq_p
        self.info = {f: getattr(info, f) for f, _ in _lib.InputInfo._fields_}
        n, nb = info.numPairs, info.numBytes
        self.sequences = np.ctypeslib.as_array(C.cast(seq_p, C.POINTER(C.c_uint8)), shape=(max(nb, 1),))[:nb]
        raw = np.ctypeslib.as_array(C.cast(pairs_p, C.POINTER(C.c_int32)), shape=(max(4 * n, 4),))[:4 * n]
        self.pairs = raw.view(PAIR_DTYPE)

    def free(self):
        if self._seq_p:
            self.sequences = self.pairs = None
            self._L.dpx_free(self._pairs_p); self._L.dpx_free(self._seq_p)
            self._pairs_p = self._seq_p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def parse_input_native(path: str) -> NativeInput:
    L = _lib.lib()
    pairs_p, seq_p, info = C.c_void_p(), C.c_void_p(), _lib.InputInfo()
    _check(L.dpx_parse_input(path.encode(), C.byref(pairs_p), C.byref(seq_p), C.byref(info)))
    return NativeInput(pairs_p, seq_p, info)


def parse_image_native(image) -> NativeInput:
    """dpx_parse_image: the parser on a file image in memory (bytes or a uint8 array)."""
    L = _lib.lib()
    a = np.ascontiguousarray(np.frombuffer(image, dtype=np.uint8) if isinstance(image, (bytes, bytearray)) else image, dtype=np.uint8)
    pairs_p, seq_p, info = C.c_void_p(), C.c_void_p(), _lib.InputInfo()
    _check(L.dpx_parse_image(a.ctypes.data, a.size, C.byref(pairs_p), C.byref(seq_p), C.byref(info)))
    return NativeInput(pairs_p, seq_p, info)


def register_input(sequences: np.ndarray, pairs: np.ndarray) -> None:
    """dpx_register_input: packs a caller-built (blob, index) once on the host; both arrays must stay alive and unmodified."""
    _check(_lib.lib().dpx_register_input(sequences.ctypes.data, sequences.size, pairs.ctypes.data, len(pairs)))


def input_sidecar(sequences: np.ndarray):
    """dpx_input_sidecar: None when `sequences` is not a registered blob, else a dict with the packed words (a copy), the word
    offsets, the symbol count, the code -> byte map, whether the lengths are uniform and whether the memory is page-locked."""
    L = _lib.lib()
    n, nw, w, off, ns, pl = C.c_size_t(), C.c_size_t(), C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    inv = (C.c_ubyte * 4)()
    kind = L.dpx_input_sidecar(sequences.ctypes.data, C.byref(n), C.byref(nw), C.byref(w), C.byref(off), C.byref(ns), C.byref(pl), inv)
    if kind == 0:
        return None
    words = np.frombuffer(C.string_at(w, 4 * nw.value), dtype=np.uint32).copy()
    woff = np.frombuffer(C.string_at(off, 4 * (n.value + 1)), dtype=np.uint32).copy()
    upload = 4 * nw.value + (0 if kind == 2 else 8 * n.value + 4)
    return dict(n_pairs=n.value, words=words, word_offsets=woff, n_symbols=ns.value, code_to_byte=bytes(inv), uniform=kind == 2,
                page_locked=bool(pl.value), upload_bytes=upload)


def unregister_input(sequences: np.ndarray) -> None:
    _lib.lib().dpx_unregister_input(sequences.ctypes.data)


def parse_fastx(path_refs: str, path_queries: str | None = None) -> ParsedInput:
    """FASTA / FASTQ records as pairs (dpx_parse_fastx): one file = records alternate reference, query; two files = paired by order."""
    L = _lib.lib()
    pairs_p, seq_p, info = C.c_void_p(), C.c_void_p(), _lib.InputInfo()
    _check(L.dpx_parse_fastx(path_refs.encode(), path_queries.encode() if path_queries else None, C.byref(pairs_p), C.byref(seq_p), C.byref(info)))
    try:
        n, nb = info.numPairs, info.numBytes
        seqs = np.ctypeslib.as_array(C.cast(seq_p, C.POINTER(C.c_uint8)), shape=(max(nb, 1),))[:nb].copy()
        pairs = np.frombuffer(C.string_at(pairs_p, n * PAIR_DTYPE.itemsize), dtype=PAIR_DTYPE).copy() if n else np.zeros(0, PAIR_DTYPE)
    finally:
        L.dpx_free(pairs_p); L.dpx_free(seq_p)
    return ParsedInput(seqs, pairs, {f: getattr(info, f) for f, _ in _lib.InputInfo._fields_})


@dataclass
class BatchResult:
    scores: np.ndarray                    # int32[n]
    end_row_col: np.ndarray | None        # int32[n,2]
    strings: list | None                  # [(REF, REL, QRY)] bytes

    def text(self, first_index: int = 0) -> bytes:
        """The reference's stdout blocks for these pairs (c++/LinearNeedlemanWunsch.cpp:207-213)."""
        parts = []
        for i, s in enumerate(self.scores):
            a, b, c = self.strings[i]
            parts.append(b"%d | %d\n" % (first_index + i, int(s)))
            parts.append(a + b"\n" + b + b"\n" + c + b"\n")
        return b"".join(parts)


def make_params(algo, match=3, mismatch=-1, gap_open=-2, gap_extend=-1, band=0, flags=OUT_SCORE) -> _lib.Params:
    return _lib.Params(algo, match, mismatch, gap_open, gap_extend, band, flags)


class Engine:
    """One device context (dpx_ctx)."""

    def __init__(self, device: int = 0):
        self.L = _lib.lib()
        self.ctx = C.c_void_p()
        _check(self.L.dpx_create(C.byref(self.ctx), device))

    def close(self):
        if self.ctx:
            self.L.dpx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        _check(self.L.dpx_set_stream(self.ctx, C.c_void_p(cuda_stream or 0)), self.ctx)

    def set_option(self, name: str, value: int):
        """dpx_set_option: debug / test knobs (the library never reads the environment)."""
        _check(self.L.dpx_set_option(self.ctx, name.encode(), int(value)), self.ctx)

    def options(self, **kw):
        """Context manager: set the given 0-default options for the duration of a `with` block, then reset them to 0."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            for k, v in kw.items():
                self.set_option(k, v)
            try:
                yield self
            finally:
                for k in kw:
                    self.set_option(k, 0)
        return cm()

    # ---- one call: host buffers in, host buffers out ---------------------------------------------
    def align_batch(self, params: _lib.Params, sequences: np.ndarray, pairs: np.ndarray) -> BatchResult:
        return _align_batch_call(self.L.dpx_align_batch, self.ctx, self.L.dpx_last_error, params, sequences, pairs, self._take_strings)

    def align_batch_text(self, params: _lib.Params, sequences: np.ndarray, pairs: np.ndarray, first_index: int = 0) -> bytes:
        """The reference's stdout blocks for the batch ("<i> | <score>\\n" + REF/REL/QRY lines), formatted on the GPU."""
        sequences = np.ascontiguousarray(sequences, dtype=np.uint8)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        txt, nb = C.c_void_p(), C.c_size_t()
        _check(self.L.dpx_align_batch_text(self.ctx, C.byref(params), sequences.ctypes.data, sequences.size, pairs.ctypes.data, len(pairs),
                                           first_index, None, None, C.byref(txt), C.byref(nb)), self.ctx)
        try:
            return C.string_at(txt, nb.value)
        finally:
            self.L.dpx_free(txt)

    def align_batch_text_all(self, params: _lib.Params, sequences: np.ndarray, pairs: np.ndarray, first_index: int = 0):
        """LinearSmithWaterman, every maximum cell walked (the reference's BACKTRACK_ALL mode): (stdout blocks, number of alignments)."""
        sequences = np.ascontiguousarray(sequences, dtype=np.uint8)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        txt, nb, na = C.c_void_p(), C.c_size_t(), C.c_longlong()
        _check(self.L.dpx_align_batch_text_all(self.ctx, C.byref(params), sequences.ctypes.data, sequences.size, pairs.ctypes.data, len(pairs),
                                               first_index, C.byref(txt), C.byref(nb), C.byref(na)), self.ctx)
        try:
            return C.string_at(txt, nb.value), na.value
        finally:
            self.L.dpx_free(txt)

    def align_file_text(self, params: _lib.Params, path: str, first_index: int = 0):
        """(text, info): the reference driver's stdout blocks for a whole input file; parser, alignment and formatting on the GPU."""
        txt, nb, info = C.c_void_p(), C.c_size_t(), _lib.InputInfo()
        _check(self.L.dpx_align_file_text(self.ctx, C.byref(params), path.encode(), first_index, C.byref(txt), C.byref(nb), C.byref(info)), self.ctx)
        try:
            return C.string_at(txt, nb.value), {f: getattr(info, f) for f, _ in _lib.InputInfo._fields_}
        finally:
            self.L.dpx_free(txt)

    def upload_image(self, image: bytes) -> "Batch":
        """Batch from a file image (3 lines per pair); the parser runs on the device."""
        return Batch(self, None, None, image=image)

    def _take_strings(self, sb, so, n):
        try:
            offs = np.frombuffer(C.string_at(so, 3 * n * C.sizeof(C.c_size_t)), dtype=np.uint64) if n else []
            base = sb.value
            out = []
            for i in range(n):
                out.append(tuple(C.string_at(base + int(offs[3 * i + k])) for k in range(3)))
            return out
        finally:
            self.L.dpx_free(sb); self.L.dpx_free(so)

    # ---- staged ----------------------------------------------------------------------------------
    def upload(self, sequences: np.ndarray, pairs: np.ndarray) -> "Batch":
        return Batch(self, sequences, pairs)

    def dpx_eval(self, op: int, a, b, c):
        a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32); c = np.ascontiguousarray(c, dtype=np.uint32)
        n = len(a)
        out = np.zeros(n, np.uint32); ph = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint8)
        _check(self.L.dpx_dpx_eval(self.ctx, op, a.ctypes.data, b.ctypes.data, c.ctypes.data, n,
                                   out.ctypes.data, ph.ctypes.data, pl.ctypes.data), self.ctx)
        return out, ph, pl

    def selftest(self) -> int:
        return self.L.dpx_selftest_dpx(self.ctx)

    def align_long_pair(self, params: _lib.Params, ref: bytes, qry: bytes):
        s = C.c_int32(); r = C.c_int64(); c = C.c_int64()
        _check(self.L.dpx_align_long_pair(self.ctx, C.byref(params), ref, len(ref), qry, len(qry),
                                          C.byref(s), C.byref(r), C.byref(c)), self.ctx)
        return s.value, r.value, c.value

    def align_long_pair_strings(self, params: _lib.Params, ref: bytes, qry: bytes):
        """One long LinearSmithWaterman pair with its alignment: ((score, end_row, end_col), (start_row, start_col),
        (REF, REL, QRY) bytes, stages) — stages = device ms of the checkpointed forward pass and of the tile fills + walk, rounds, tiles walked, tile shape."""
        s = C.c_int32(); r = C.c_int64(); c = C.c_int64(); r0 = C.c_int64(); c0 = C.c_int64()
        blob = C.c_void_p(); n = C.c_size_t(); ms = (C.c_double * 6)()
        _check(self.L.dpx_align_long_pair_strings(self.ctx, C.byref(params), ref, len(ref), qry, len(qry), C.byref(s), C.byref(r), C.byref(c),
                                                  C.byref(r0), C.byref(c0), C.byref(blob), C.byref(n), ms), self.ctx)
        L = n.value
        raw = C.string_at(blob.value, 3 * (L + 1))
        self.L.dpx_free(blob)
        lines = tuple(raw[k * (L + 1): k * (L + 1) + L] for k in range(3))
        return (s.value, r.value, c.value), (r0.value, c0.value), lines, dict(fwd_ms=ms[0], rounds=int(ms[1]), walk_ms=ms[2], tiles=int(ms[3]), tile_rows=int(ms[4]), tile_cols=int(ms[5]))


def _align_batch_call(fn, handle, errfn, params, sequences, pairs, take_strings) -> BatchResult:
    n = len(pairs)
    sequences = np.ascontiguousarray(sequences, dtype=np.uint8)
    pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
    scores = np.zeros(n, dtype=np.int32)
    end_rc = np.zeros((n, 2), dtype=np.int32)
    sb, so = C.c_void_p(), C.c_void_p()
    want = bool(params.flags & OUT_STRINGS)
    st = fn(handle, C.byref(params), sequences.ctypes.data, sequences.size, pairs.ctypes.data, n, scores.ctypes.data, end_rc.ctypes.data,
            C.byref(sb) if want else None, C.byref(so) if want else None)
    if st != 0:
        raise DpxError(st, errfn(handle).decode())
    return BatchResult(scores, end_rc, take_strings(sb, so, n) if want else None)


class MultiEngine:
    """Several GPUs of one node behind ONE host process (dpx_create_multi): contiguous shards balanced by cell count, one worker
    thread + context per device, results in pair order (multi-GPU mode A through the C ABI)."""

    def __init__(self, devices=None, n_devices: int | None = None):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(self.L.dpx_create_multi(C.byref(self.h), arr, len(devices)))
        else:
            _check(self.L.dpx_create_multi(C.byref(self.h), None, int(n_devices or self.L.dpx_device_count())))

    @property
    def n_devices(self) -> int:
        return self.L.dpx_multi_device_count(self.h)

    def set_option(self, name: str, value: int):
        st = self.L.dpx_multi_set_option(self.h, name.encode(), int(value))
        if st:
            raise DpxError(st, self.L.dpx_multi_last_error(self.h).decode())

    def shard_bounds(self, pairs: np.ndarray, n_shards: int | None = None):
        k = n_shards or self.n_devices
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        out = (C.c_size_t * (k + 1))()
        _check(self.L.dpx_multi_shard_bounds(pairs.ctypes.data, len(pairs), k, out))
        return list(out)

    def align_batch(self, params: _lib.Params, sequences: np.ndarray, pairs: np.ndarray) -> BatchResult:
        return _align_batch_call(self.L.dpx_multi_align_batch, self.h, self.L.dpx_multi_last_error, params, sequences, pairs,
                                 lambda sb, so, n: Engine._take_strings(self, sb, so, n))

    def align_batch_text(self, params: _lib.Params, sequences: np.ndarray, pairs: np.ndarray, first_index: int = 0) -> bytes:
        sequences = np.ascontiguousarray(sequences, dtype=np.uint8)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        txt, nb = C.c_void_p(), C.c_size_t()
        st = self.L.dpx_multi_align_batch_text(self.h, C.byref(params), sequences.ctypes.data, sequences.size, pairs.ctypes.data, len(pairs),
                                               first_index, None, None, C.byref(txt), C.byref(nb))
        if st:
            raise DpxError(st, self.L.dpx_multi_last_error(self.h).decode())
        try:
            return C.string_at(txt, nb.value)
        finally:
            self.L.dpx_free(txt)

    def close(self):
        if self.h:
            self.L.dpx_destroy_multi(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Pairs resident in HBM (dpx_batch): upload once, run any number of times, fetch."""

    def __init__(self, eng: Engine, sequences, pairs, image: bytes | None = None):
        self.eng = eng
        self.h = C.c_void_p()
        self.params = None
        if image is not None:
            info = _lib.InputInfo()
            _check(eng.L.dpx_batch_upload_image(eng.ctx, image, len(image), C.byref(self.h), C.byref(info)), eng.ctx)
            self.n = info.numPairs
            self.info = {f: getattr(info, f) for f, _ in _lib.InputInfo._fields_}
            return
        self.n = len(pairs)
        self._seq = np.ascontiguousarray(sequences, dtype=np.uint8)
        self._pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        _check(eng.L.dpx_batch_upload(eng.ctx, self._seq.ctypes.data, self._seq.size, self._pairs.ctypes.data, self.n,
                                      C.byref(self.h)), eng.ctx)

    def run(self, params: _lib.Params):
        self.params = params
        _check(self.eng.L.dpx_batch_run(self.h, C.byref(params)), self.eng.ctx)

    def sync(self):
        _check(self.eng.L.dpx_batch_sync(self.h), self.eng.ctx)

    def stats(self) -> dict:
        s = _lib.RunStats()
        _check(self.eng.L.dpx_batch_stats(self.h, C.byref(s)), self.eng.ctx)
        return {f: getattr(s, f) for f, _ in _lib.RunStats._fields_}

    def fetch(self, scores: np.ndarray | None = None, end_rc: np.ndarray | None = None) -> BatchResult:
        n = self.n
        scores = np.zeros(n, dtype=np.int32) if scores is None else scores
        end_rc = np.zeros((n, 2), dtype=np.int32) if end_rc is None else end_rc
        sb, so = C.c_void_p(), C.c_void_p()
        want = bool(self.params.flags & OUT_STRINGS)
        _check(self.eng.L.dpx_batch_fetch(self.h, scores.ctypes.data, end_rc.ctypes.data,
                                          C.byref(sb) if want else None, C.byref(so) if want else None), self.eng.ctx)
        return BatchResult(scores, end_rc, self.eng._take_strings(sb, so, n) if want else None)

    def free(self):
        if self.h:
            self.eng.L.dpx_batch_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- the reference's per-pair classes -----------------------------------------------------------------
_default_engine = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _as_bytes(s) -> bytes:
    return s.encode() if isinstance(s, str) else bytes(s)


class SequenceAligner:
    """c++/SequenceAligner.h:6-28: holds reference_str, query_str, pairNum; subclasses implement align()."""
    _algo = None

    def __init__(self, input_reference, input_query, pairNum: int):
        self.reference_str = _as_bytes(input_reference)
        self.query_str = _as_bytes(input_query)
        self.pairNum = int(pairNum)
        self.score = None
        self.end_row_col = None
        self.reference_sequence = self.pair_relation = self.query_sequence = None
        self.out = sys.stdout

    def _params(self) -> _lib.Params:
        raise NotImplementedError

    # the reference's six virtuals; on the GPU the first three are one fused batch call
    def init_matrix(self): pass
    def print_matrix(self): raise NotImplementedError("matrices never leave the GPU")
    def score_matrix(self): self._run()
    def backtrack(self):
        if self.score is None: self._run()
    def print_results(self):
        w = getattr(self.out, "buffer", self.out)
        w.write(self.text())

    def _run(self):
        r, q = self.reference_str, self.query_str
        blob = np.frombuffer(r + b"\0" + q + b"\0", dtype=np.uint8)
        pairs = np.array([(0, len(r), len(r) + 1, len(q))], dtype=PAIR_DTYPE)
        res = default_engine().align_batch(self._params(), blob, pairs)
        self.score = int(res.scores[0])
        self.end_row_col = (int(res.end_row_col[0][0]), int(res.end_row_col[0][1]))
        self.reference_sequence, self.pair_relation, self.query_sequence = res.strings[0]

    def text(self) -> bytes:
        return b"%d | %d\n" % (self.pairNum, self.score) + self.reference_sequence + b"\n" + self.pair_relation + b"\n" + self.query_sequence + b"\n"

    def align(self):
        self.init_matrix(); self.score_matrix(); self.backtrack(); self.print_results()


class LinearNeedlemanWunsch(SequenceAligner):
    def __init__(self, input_reference, input_query, pairNum, match_weight, mismatch_weight, gap_weight):
        super().__init__(input_reference, input_query, pairNum)
        self.w = (match_weight, mismatch_weight, gap_weight)

    def _params(self):
        return make_params(LNW, self.w[0], self.w[1], self.w[2], 0, 0, OUT_SCORE | OUT_END_COORDS | OUT_STRINGS)


class LinearSmithWaterman(SequenceAligner):
    def __init__(self, input_reference, input_query, pairNum, match_weight, mismatch_weight, gap_weight):
        super().__init__(input_reference, input_query, pairNum)
        self.w = (match_weight, mismatch_weight, gap_weight)

    def _params(self):
        return make_params(LSW, self.w[0], self.w[1], self.w[2], 0, 0, OUT_SCORE | OUT_END_COORDS | OUT_STRINGS)


class AffineNeedlemanWunsch(SequenceAligner):
    def __init__(self, input_reference, input_query, pairNum, matchWeight, mismatchWeight, gapOpenWeight, gapExtendWeight):
        super().__init__(input_reference, input_query, pairNum)
        self.w = (matchWeight, mismatchWeight, gapOpenWeight, gapExtendWeight)

    def _params(self):
        return make_params(ANW, self.w[0], self.w[1], self.w[2], self.w[3], 0, OUT_SCORE | OUT_END_COORDS | OUT_STRINGS)


class BandedSmithWaterman(SequenceAligner):
    def __init__(self, input_reference, input_query, match_weight, mismatch_weight, gap_weight, pairNum, band_width=64):
        super().__init__(input_reference, input_query, pairNum)
        self.w = (match_weight, mismatch_weight, gap_weight)
        self.band_width = int(band_width)

    def _params(self):
        return make_params(BSW, self.w[0], self.w[1], self.w[2], 0, self.band_width, OUT_SCORE | OUT_END_COORDS | OUT_STRINGS)


class AffineBandedSmithWaterman(SequenceAligner):
    """Not a reference class: affine banded Smith-Waterman (DPX_ALGO_ABSW, include/dpxalign.h), the variant the reference names
    as a TODO (python/LinearBandedSmithWaterman.py:8).  Constructor shaped like BandedSmithWaterman's (pairNum after the weights)."""

    def __init__(self, input_reference, input_query, match_weight, mismatch_weight, gap_open_weight, gap_extend_weight, pairNum, band_width=64):
        super().__init__(input_reference, input_query, pairNum)
        self.w = (match_weight, mismatch_weight, gap_open_weight, gap_extend_weight)
        self.band_width = int(band_width)

    def _params(self):
        return make_params(ABSW, self.w[0], self.w[1], self.w[2], self.w[3], self.band_width, OUT_SCORE | OUT_END_COORDS | OUT_STRINGS)
