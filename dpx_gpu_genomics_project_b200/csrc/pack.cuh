// pack.cuh — alphabet scan and 2-bit packing of the parseInput blob on the device.
//
// The reference aligners compare raw bytes (c++/LinearNeedlemanWunsch.cpp:108), so any byte->code map that
// is injective on the bytes actually present preserves every result.  present_kernel builds the 256-bit
// presence set of all sequence bytes of the batch; the host ranks the present bytes (code = rank) and, when
// there are at most four symbols, pack2_kernel writes 2-bit codes, 16 bases per 32-bit word.  A fifth symbol
// (the data sets use '0'..'4', reference correct-outputs/LNW/web-scraper-LNW.py:5-12) is the escape: the
// batch then stays on the byte-compare kernels.
//
// Packed layout: pair p owns words [pk_off[p], pk_off[p+1]) : ceil(R/16) reference words followed by
// ceil(Q/16) query words; base k of a sequence sits in word k/16 at bits 2*(k%16).
#pragma once
#include "common.cuh"

namespace dpx {

__global__ void __launch_bounds__(256) present_kernel(const uint8_t* __restrict__ blob, const dpx_seq_pair* __restrict__ pairs,
                                                       int n_pairs, uint32_t* __restrict__ present /*[8]*/) {
    __shared__ uint32_t sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = warp; p < n_pairs; p += nwarps) {
        const dpx_seq_pair pr = pairs[p];
        for (int k = lane; k < pr.referenceSize; k += 32) { const uint8_t c = blob[pr.referenceIdx + k]; loc[c >> 5] |= 1u << (c & 31); }
        for (int k = lane; k < pr.querySize; k += 32)     { const uint8_t c = blob[pr.queryIdx + k];     loc[c >> 5] |= 1u << (c & 31); }
    }
    #pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t v = __reduce_or_sync(0xffffffffu, loc[w]);
        if (lane == 0 && v) atomicOr(&sh[w], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicOr(&present[threadIdx.x], sh[threadIdx.x]);
}

struct PackLut { uint8_t code[256]; };

__global__ void __launch_bounds__(256) pack2_kernel(const uint8_t* __restrict__ blob, const dpx_seq_pair* __restrict__ pairs,
                                                     int n_pairs, const unsigned long long* __restrict__ pk_off,
                                                     uint32_t* __restrict__ packed, const PackLut lut) {
    __shared__ uint8_t code[256];
    code[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < n_pairs; p += nwarps) {
        const dpx_seq_pair pr = pairs[p];
        uint32_t* __restrict__ out = packed + pk_off[p];
        const int rw = (pr.referenceSize + 15) >> 4, qw = (pr.querySize + 15) >> 4;
        for (int w = lane; w < rw + qw; w += 32) {
            const bool isq = w >= rw;
            const uint8_t* __restrict__ src = blob + (isq ? pr.queryIdx : pr.referenceIdx);
            const int len = isq ? pr.querySize : pr.referenceSize;
            const int k0 = (isq ? w - rw : w) << 4;
            uint32_t v = 0;
            #pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k0 + k < len) v |= (uint32_t)(code[src[k0 + k]] & 3u) << (2 * k);
            out[w] = v;
        }
    }
}

}  // namespace dpx
