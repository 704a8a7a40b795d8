"""Multi-GPU mode B: ONE very long pair, the reference columns split into one stripe per GPU.

Stripe g holds columns [bounds[g], bounds[g+1]) and is a chain of warps (csrc/longpair.cuh); the right edge of the last
warp of stripe g streams, row by row, into the inbox ring of stripe g+1 — peer memory written with plain NVLink P2P
stores + st.release.sys, the consumer polls its own HBM.  The stripes therefore run pipelined along the anti-diagonal:
stripe g+1 trails stripe g by the depth of stripe g's warp chain.  No NCCL collective is on the data path; torch.distributed
is only used to pass the 64-byte CUDA-IPC handles around, for the start barrier and to gather the 8 result tuples.
The reference has no multi-GPU code and cannot even allocate this problem (8 B per cell).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import DpxError, Engine


def stripe_bounds(R: int, n: int, align: int = 1024) -> list[int]:
    """Column boundaries of n stripes over R columns: equal widths rounded to `align` (whole warps at up to 32 columns per lane;
    the library narrows the lanes of a stripe whose width is not a multiple of 32 K), last stripe takes the rest."""
    if n <= 1:
        return [0, R]
    w = -(-R // n)
    w = -(-w // align) * align
    b = [min(R, g * w) for g in range(n)] + [R]
    return b


def better(a, b):
    """Reference end-cell order between two (score, row, col) candidates: higher score, then smaller row, then smaller column
    (first strict maximum in row-major order, c++/LinearSmithWaterman.cpp:145-157)."""
    if b[0] > a[0] or (b[0] == a[0] and b[0] > 0 and (b[1] < a[1] or (b[1] == a[1] and b[2] < a[2]))):
        return b
    return a


def reduce_results(results):
    best = (0, 0, 0)
    for r in results:
        if r is not None and r[0] > 0:
            best = better(best, tuple(r)) if best[0] > 0 else tuple(r)
    return best


class Stripe:
    def __init__(self, eng: Engine, params, ref_stripe: bytes, col_offset: int, qry: bytes, index: int, n: int):
        self.eng = eng
        self.h = C.c_void_p()
        st = eng.L.dpx_stripe_create(eng.ctx, C.byref(params), ref_stripe, len(ref_stripe), col_offset, qry, len(qry), index, n, C.byref(self.h))
        if st:
            raise DpxError(st, eng.L.dpx_last_error(eng.ctx).decode())

    def _chk(self, st):
        if st:
            raise DpxError(st, self.eng.L.dpx_last_error(self.eng.ctx).decode())

    def export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._chk(self.eng.L.dpx_stripe_export(self.h, buf))
        return buf.raw

    def connect(self, prev_handle: bytes | None, next_handle: bytes | None):
        self._chk(self.eng.L.dpx_stripe_connect(self.h, prev_handle, next_handle))

    def reset(self):
        self._chk(self.eng.L.dpx_stripe_reset(self.h))

    def run(self):
        self._chk(self.eng.L.dpx_stripe_run(self.h))

    def result(self):
        s = C.c_int32(); r = C.c_int64(); c = C.c_int64(); ms = C.c_double()
        self._chk(self.eng.L.dpx_stripe_result(self.h, C.byref(s), C.byref(r), C.byref(c), C.byref(ms)))
        return (s.value, r.value, c.value), ms.value

    def free(self):
        if self.h:
            self.eng.L.dpx_stripe_free(self.h)
            self.h = C.c_void_p()


class StripedLongPair:
    """One rank's part of a striped long-pair alignment.  Usage (every rank):
        job = StripedLongPair(eng, params, ref, qry, rank, world, dist);  res, ms = job.run()   # res on every rank"""

    def __init__(self, eng: Engine, params, ref: bytes, qry: bytes, rank: int, world: int, dist):
        self.rank, self.world, self.dist = rank, world, dist
        self.bounds = stripe_bounds(len(ref), world)
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        self.empty = hi <= lo
        self.stripe = None if self.empty else Stripe(eng, params, ref[lo:hi], lo, qry, rank, self._n_active())
        handles = [None] * world
        mine = None if self.empty else self.stripe.export()
        if world > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles = [mine]
        if not self.empty:
            prev_h = handles[rank - 1] if rank > 0 else None
            next_h = handles[rank + 1] if rank + 1 < world and handles[rank + 1] is not None else None
            self.stripe.connect(prev_h, next_h)

    def _n_active(self) -> int:
        return sum(1 for g in range(self.world) if self.bounds[g + 1] > self.bounds[g])

    def run(self):
        """Returns ((score, row, col), max-over-ranks kernel ms).  Kernels of all ranks run concurrently (one per GPU)."""
        import torch
        if not self.empty:
            self.stripe.reset()
        if self.world > 1:
            self.dist.barrier()
        if not self.empty:
            self.stripe.run()
            res, ms = self.stripe.result()
        else:
            res, ms = None, 0.0
        if self.world > 1:
            allres = [None] * self.world
            self.dist.all_gather_object(allres, (res, ms))
            return reduce_results([r for r, _ in allres]), max(m for _, m in allres)
        return reduce_results([res]), ms

    def free(self):
        if self.stripe:
            self.stripe.free()
