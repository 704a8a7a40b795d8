"""dpx_gpu_genomics_project_b200 — B200 (sm_100a) pairwise-alignment engine.

Drop-in for the alignment path of mickgordinier/DPX_GPU_Genomics_Project: same SequenceAligner
subclasses, parseInput format and backtrack output bytes, computed by hand-written CUDA kernels behind
the C ABI in include/dpxalign.h (libdpxalign.so).  No CPU fallback.
"""
from .api import (ANW, BSW, LNW, LSW, OUT_END_COORDS, OUT_SCORE, OUT_STRINGS, PAIR_DTYPE,  # noqa: F401
                  AffineNeedlemanWunsch, BandedSmithWaterman, Batch, BatchResult, DpxError, Engine,
                  LinearNeedlemanWunsch, LinearSmithWaterman, ParsedInput, SequenceAligner,
                  make_params, parse_input)
