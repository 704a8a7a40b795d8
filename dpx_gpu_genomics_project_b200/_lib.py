"""Loader for libdpxalign.so (the C ABI in include/dpxalign.h) through ctypes.

The shared object is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing this module raises, and every alignment call needs a CUDA
device (``dpx_create`` returns DPX_ERR_NO_DEVICE otherwise).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdpxalign.so")
HEADER = os.path.join(ROOT, "include", "dpxalign.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-shared"]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources() + [HEADER])


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libdpxalign.so (in-tree)."""
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              [os.path.join(CSRC, "dpxalign.cu"), os.path.join(CSRC, "host_pack.cpp"), os.path.join(CSRC, "host_multi.cpp"),
               "-Xcompiler", "-pthread", "-o", LIB_PATH]
        subprocess.run(cmd, check=True)
    return LIB_PATH


class Params(C.Structure):
    _fields_ = [("algo", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("band", C.c_int32), ("flags", C.c_uint32)]


class InputInfo(C.Structure):
    _fields_ = [("numPairs", C.c_size_t), ("numBytes", C.c_size_t), ("numCells", C.c_size_t),
                ("minReferenceLength", C.c_size_t), ("minQueryLength", C.c_size_t),
                ("maxReferenceLength", C.c_size_t), ("maxQueryLength", C.c_size_t),
                ("avgReferenceLength", C.c_double), ("avgQueryLength", C.c_double)]


class RunStats(C.Structure):
    _fields_ = [("fill_ms", C.c_double), ("backtrack_ms", C.c_double), ("total_ms", C.c_double),
                ("cells", C.c_uint64), ("traceback_bytes", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("kernel_id", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(libdpxalign has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    L.dpx_abi_version.restype = C.c_int
    L.dpx_strerror.restype = C.c_char_p; L.dpx_strerror.argtypes = [C.c_int]
    L.dpx_device_count.restype = C.c_int
    L.dpx_create.restype = C.c_int; L.dpx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.dpx_destroy.restype = None; L.dpx_destroy.argtypes = [vp]
    L.dpx_last_error.restype = C.c_char_p; L.dpx_last_error.argtypes = [vp]
    L.dpx_set_stream.restype = C.c_int; L.dpx_set_stream.argtypes = [vp, vp]
    L.dpx_set_option.restype = C.c_int; L.dpx_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.dpx_trim.restype = None; L.dpx_trim.argtypes = []
    L.dpx_parse_image.restype = C.c_int
    L.dpx_parse_image.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), C.POINTER(InputInfo)]
    L.dpx_register_input.restype = C.c_int; L.dpx_register_input.argtypes = [vp, C.c_size_t, vp, C.c_size_t]
    L.dpx_input_sidecar.restype = C.c_int
    L.dpx_input_sidecar.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
    L.dpx_unregister_input.restype = None; L.dpx_unregister_input.argtypes = [vp]
    L.dpx_bind_host_to_device.restype = C.c_int; L.dpx_bind_host_to_device.argtypes = [C.c_int]
    L.dpx_create_multi.restype = C.c_int; L.dpx_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.dpx_destroy_multi.restype = None; L.dpx_destroy_multi.argtypes = [vp]
    L.dpx_multi_last_error.restype = C.c_char_p; L.dpx_multi_last_error.argtypes = [vp]
    L.dpx_multi_device_count.restype = C.c_int; L.dpx_multi_device_count.argtypes = [vp]
    L.dpx_multi_set_option.restype = C.c_int; L.dpx_multi_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.dpx_multi_align_batch.restype = C.c_int
    L.dpx_multi_align_batch.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.dpx_multi_align_batch_text.restype = C.c_int
    L.dpx_multi_align_batch_text.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, C.c_longlong, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.dpx_multi_shard_bounds.restype = C.c_int
    L.dpx_multi_shard_bounds.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
    L.dpx_parse_input.restype = C.c_int
    L.dpx_parse_input.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(InputInfo)]
    L.dpx_parse_fastx.restype = C.c_int
    L.dpx_parse_fastx.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(InputInfo)]
    L.dpx_free.restype = None; L.dpx_free.argtypes = [vp]
    L.dpx_align_batch.restype = C.c_int
    L.dpx_align_batch.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.dpx_batch_upload.restype = C.c_int
    L.dpx_batch_upload.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]
    L.dpx_batch_run.restype = C.c_int; L.dpx_batch_run.argtypes = [vp, C.POINTER(Params)]
    L.dpx_batch_sync.restype = C.c_int; L.dpx_batch_sync.argtypes = [vp]
    L.dpx_batch_fetch.restype = C.c_int; L.dpx_batch_fetch.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.dpx_batch_free.restype = None; L.dpx_batch_free.argtypes = [vp]
    L.dpx_batch_fetch_text.restype = C.c_int; L.dpx_batch_fetch_text.argtypes = [vp, C.c_longlong, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.dpx_align_batch_text.restype = C.c_int
    L.dpx_align_batch_text.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, C.c_longlong, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.dpx_batch_upload_image.restype = C.c_int; L.dpx_batch_upload_image.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp), C.POINTER(InputInfo)]
    L.dpx_align_file_text.restype = C.c_int
    L.dpx_align_file_text.argtypes = [vp, C.POINTER(Params), C.c_char_p, C.c_longlong, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(InputInfo)]
    L.dpx_batch_stats.restype = C.c_int; L.dpx_batch_stats.argtypes = [vp, C.POINTER(RunStats)]
    L.dpx_align_long_pair.restype = C.c_int
    L.dpx_align_long_pair.argtypes = [vp, C.POINTER(Params), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                      i32p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.dpx_align_batch_text_all.restype = C.c_int
    L.dpx_align_batch_text_all.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, C.c_longlong, C.POINTER(vp), C.POINTER(C.c_size_t),
                                           C.POINTER(C.c_longlong)]
    L.dpx_align_long_pair_strings.restype = C.c_int
    L.dpx_align_long_pair_strings.argtypes = [vp, C.POINTER(Params), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                              i32p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                              C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
    L.dpx_stripe_create.restype = C.c_int
    L.dpx_stripe_create.argtypes = [vp, C.POINTER(Params), C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp)]
    L.dpx_stripe_export.restype = C.c_int; L.dpx_stripe_export.argtypes = [vp, vp]
    L.dpx_stripe_connect.restype = C.c_int; L.dpx_stripe_connect.argtypes = [vp, vp, vp]
    L.dpx_stripe_reset.restype = C.c_int; L.dpx_stripe_reset.argtypes = [vp]
    L.dpx_stripe_run.restype = C.c_int; L.dpx_stripe_run.argtypes = [vp]
    L.dpx_stripe_result.restype = C.c_int
    L.dpx_stripe_result.argtypes = [vp, i32p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.dpx_stripe_free.restype = None; L.dpx_stripe_free.argtypes = [vp]
    L.dpx_selftest_dpx.restype = C.c_int; L.dpx_selftest_dpx.argtypes = [vp]
    L.dpx_dpx_eval.restype = C.c_int
    L.dpx_dpx_eval.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp]
    _lib = L
    return L
