// common.cuh — shared host/device definitions of libdpxalign (sm_100a only).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include "../../include/dpxalign.h"

#ifndef __CUDA_ARCH__
#define DPX_HD __host__ __device__ inline
#else
#define DPX_HD __host__ __device__ __forceinline__
#endif

namespace dpx {

// 2-bit direction codes shared by every fill kernel and the backtrack kernels.
//   NW / Gotoh H : DIAG, UP (QUERY_DELETION), LEFT (QUERY_INSERTION)   reference c++/backtrack.h:14-20
//   SW           : STOP means H == 0 (the reference stops when memo[next] == 0, c++/LinearSmithWaterman.cpp:222)
// Gotoh adds bit2 = D came from GAP_OPEN, bit3 = I came from GAP_OPEN (c++/backtrack.h:23-27).
enum : uint32_t { C_STOP = 0, C_DIAG = 1, C_UP = 2, C_LEFT = 3, C_DOPEN = 4, C_IOPEN = 8 };

constexpr int32_t NEG_INF = -(1 << 29);   // "minus infinity" that survives a few additions in int32

// Geometry of the warp-wavefront traceback slab (wavefront.cuh writes it, backtrack.cuh reads it).
// A stripe is 32*K query rows; lane t owns rows [t*K, t*K+K) of the stripe and at step s works on
// column jstart + s - t.  Every lane packs K codes of CB bits per step, SPW steps per 32-bit word;
// words are stored [stripe][step/SPW][lane] so that one warp-step store is one 128-byte line.
struct WfGeom {
    int K, CB, SPW;          // rows per lane, code bits, steps per word
    int Q, R, band;          // band < 0: unbanded
    int rows_per_stripe;     // 32*K
    int nstripes;
    int max_cols;            // widest column range of any stripe
    int ngroups;             // words per lane per stripe = ceil((max_cols + 31) / SPW)

    DPX_HD static WfGeom make(int K, int CB, int Q, int R, int band) {
        WfGeom g;
        g.K = K; g.CB = CB; g.SPW = 32 / (K * CB); g.Q = Q; g.R = R; g.band = band;
        g.rows_per_stripe = 32 * K;
        g.nstripes = (Q + g.rows_per_stripe - 1) / g.rows_per_stripe;
        int mc = R;
        if (band >= 0) { long long w = (long long)g.rows_per_stripe + 2LL * band; if (w < mc) mc = (int)w; }
        g.max_cols = mc < 0 ? 0 : mc;
        g.ngroups = (g.max_cols + 31 + g.SPW - 1) / g.SPW;
        return g;
    }
    DPX_HD int jstart(int s) const {
        if (band < 0) return 1;
        long long v = (long long)s * rows_per_stripe + 1 - band;
        return v < 1 ? 1 : (int)v;
    }
    DPX_HD int jend(int s) const {
        if (band < 0) return R;
        long long v = (long long)(s + 1) * rows_per_stripe + band;
        return v > R ? R : (int)v;
    }
    DPX_HD unsigned long long words() const { return (unsigned long long)nstripes * ngroups * 32ull; }
};

}  // namespace dpx
