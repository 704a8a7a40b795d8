// pack.cuh — alphabet scan and 2-bit packing of the parseInput blob on the device.
//
// The reference aligners compare raw bytes (c++/LinearNeedlemanWunsch.cpp:108), so any byte->code map that
// is injective on the bytes actually present preserves every result.  present_kernel builds the 256-bit
// presence set of all sequence bytes of the batch; the host ranks the present bytes (code = rank) and, when
// there are at most four symbols, pack2_kernel writes 2-bit codes, 16 bases per 32-bit word.  A fifth symbol
// (the data sets use '0'..'4', reference correct-outputs/LNW/web-scraper-LNW.py:5-12) is the escape: the
// batch then stays on the byte-compare kernels.
//
// Packed layout: pair p starts at word pk_off[p] (or p * pk_stride when all pairs have equal lengths): ceil(R/16) reference words followed by
// ceil(Q/16) query words; base k of a sequence sits in word k/16 at bits 2*(k%16).
#pragma once
#include "common.cuh"

namespace dpx {

// Per-batch facts the host needs before it can choose kernels, gathered in ONE pass on the device so the host
// never loops over the pairs: alphabet presence set, length extrema, cell count, packed size, string-slot
// size, and a validity flag for the index (every range inside the blob).
struct BatchInfo {
    uint32_t present[8];
    int max_r, max_q;
    int min_r_inv, min_q_inv;              // INT_MAX - min (so that a zeroed struct works with atomicMax)
    unsigned long long cells;              // sum Q*R
    unsigned long long packed_words;       // sum ceil(R/16) + ceil(Q/16)
    unsigned long long str_bytes;          // sum 3*(Q+R+1)
    int invalid;
    int pad;
    unsigned long long sum_r, sum_q;       // sum of reference / query lengths (inputInfo averages)
};

__global__ void __launch_bounds__(256) prep_kernel(const uint8_t* __restrict__ blob, long long byte_lo, long long byte_hi,
                                                   const dpx_seq_pair* __restrict__ pairs, int n_pairs, BatchInfo* __restrict__ info,
                                                   unsigned long long* __restrict__ pk_words, unsigned long long* __restrict__ str_len) {
    __shared__ uint32_t sh[8];
    __shared__ unsigned long long sh_cells, sh_words, sh_str, sh_sumr, sh_sumq;
    __shared__ int sh_maxr, sh_maxq, sh_minr, sh_minq, sh_bad;
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sh_cells = 0; sh_words = 0; sh_str = 0; sh_sumr = 0; sh_sumq = 0; sh_maxr = 0; sh_maxq = 0; sh_minr = 0; sh_minq = 0; sh_bad = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long cells = 0, words = 0, strb = 0, sumr = 0, sumq = 0;
    int maxr = 0, maxq = 0, minr = 0, minq = 0, bad = 0;
    for (int p = warp; p < n_pairs; p += nwarps) {
        const dpx_seq_pair pr = pairs[p];
        const bool ok = pr.referenceSize >= 0 && pr.querySize >= 0 && pr.referenceIdx >= byte_lo && pr.queryIdx >= byte_lo &&
                        (long long)pr.referenceIdx + pr.referenceSize <= byte_hi && (long long)pr.queryIdx + pr.querySize <= byte_hi;
        if (!ok) { bad = 1; if (lane == 0) { if (pk_words) pk_words[p] = 0; if (str_len) str_len[p] = 0; } continue; }
        for (int k = lane; k < pr.referenceSize; k += 32) { const uint8_t c = blob[pr.referenceIdx + k]; loc[c >> 5] |= 1u << (c & 31); }
        for (int k = lane; k < pr.querySize; k += 32)     { const uint8_t c = blob[pr.queryIdx + k];     loc[c >> 5] |= 1u << (c & 31); }
        if (lane == 0) {
            const unsigned long long w = (unsigned long long)((pr.referenceSize + 15) >> 4) + (unsigned long long)((pr.querySize + 15) >> 4);
            const unsigned long long sl = 3ull * ((unsigned long long)pr.referenceSize + pr.querySize + 1);
            if (pk_words) pk_words[p] = w;
            if (str_len) str_len[p] = sl;
            cells += (unsigned long long)pr.referenceSize * (unsigned long long)pr.querySize; words += w; strb += sl;
            sumr += (unsigned long long)pr.referenceSize; sumq += (unsigned long long)pr.querySize;
            maxr = max(maxr, pr.referenceSize); maxq = max(maxq, pr.querySize);
            minr = max(minr, 0x7fffffff - pr.referenceSize); minq = max(minq, 0x7fffffff - pr.querySize);
        }
    }
    #pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t v = __reduce_or_sync(0xffffffffu, loc[w]);
        if (lane == 0 && v) atomicOr(&sh[w], v);
    }
    if (__any_sync(0xffffffffu, bad)) { if (lane == 0) atomicOr(&sh_bad, 1); }
    if (lane == 0) {
        atomicAdd(&sh_cells, cells); atomicAdd(&sh_words, words); atomicAdd(&sh_str, strb); atomicAdd(&sh_sumr, sumr); atomicAdd(&sh_sumq, sumq);
        atomicMax(&sh_maxr, maxr); atomicMax(&sh_maxq, maxq); atomicMax(&sh_minr, minr); atomicMax(&sh_minq, minq);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicOr(&info->present[threadIdx.x], sh[threadIdx.x]);
    if (threadIdx.x == 0) {
        atomicAdd(&info->cells, sh_cells); atomicAdd(&info->packed_words, sh_words); atomicAdd(&info->str_bytes, sh_str);
        atomicAdd(&info->sum_r, sh_sumr); atomicAdd(&info->sum_q, sh_sumq);
        atomicMax(&info->max_r, sh_maxr); atomicMax(&info->max_q, sh_maxq);
        atomicMax(&info->min_r_inv, sh_minr); atomicMax(&info->min_q_inv, sh_minq);
        if (sh_bad) atomicOr(&info->invalid, 1);
    }
}

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

// ---- parser on the device (replaces the byte loop of c++/parseInput.cpp:78-113) ---------------------------------------
// The file image is uploaded as is; a stream compaction lists the newline positions, three consecutive newlines make a pair:
//   header \n REF \n QRY \n    ->  referenceIdx = nl[3k] + 1, referenceSize = nl[3k+1] - referenceIdx, queryIdx = nl[3k+1] + 1, ...
struct IsNewline {
    const uint8_t* blob;
    __device__ __forceinline__ bool operator()(int i) const { return blob[i] == (uint8_t)'\n'; }
};

__global__ void __launch_bounds__(256) count_newlines_kernel(const uint8_t* __restrict__ blob, long long n, int* __restrict__ count) {
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += (blob[i] == (uint8_t)'\n');
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void __launch_bounds__(256) pairs_from_newlines_kernel(const int* __restrict__ nl, int n_pairs, dpx_seq_pair* __restrict__ pairs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const int h = nl[3 * k], r = nl[3 * k + 1], q = nl[3 * k + 2];
    dpx_seq_pair p;
    p.referenceIdx = h + 1; p.referenceSize = r - (h + 1);
    p.queryIdx = r + 1;     p.querySize = q - (r + 1);
    pairs[k] = p;
}

// seqPair index of a fixed-length file rebuilt on the device: pair k = first + k * stride (same sizes, same query offset)
__global__ void __launch_bounds__(256) regular_pairs_kernel(dpx_seq_pair* __restrict__ pairs, int n, const dpx_seq_pair first, int stride) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    dpx_seq_pair p = first;
    p.referenceIdx += k * stride; p.queryIdx += k * stride;
    pairs[k] = p;
}

struct PackLut { uint8_t code[256]; };

// byte -> code (0..7) for every byte of the blob, same indexing as the blob: alphabets of 5..8 symbols
__global__ void __launch_bounds__(256) code_bytes_kernel(const uint8_t* __restrict__ blob, long long n, uint8_t* __restrict__ codes, const PackLut lut) {
    __shared__ uint8_t code[256];
    code[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) codes[i] = code[blob[i]] & 7u;
}

__global__ void __launch_bounds__(256) pack2_kernel(const uint8_t* __restrict__ blob, const dpx_seq_pair* __restrict__ pairs,
                                                     int n_pairs, const unsigned long long* __restrict__ pk_off,
                                                     unsigned long long pk_stride,      // used when pk_off == nullptr (uniform lengths)
                                                     uint32_t* __restrict__ packed, const PackLut lut,
                                                     int* __restrict__ unknown_symbol /* nullable: set to 1 when a byte has code 0xFF */) {
    __shared__ uint8_t code[256];
    code[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < n_pairs; p += nwarps) {
        const dpx_seq_pair pr = pairs[p];
        uint32_t* __restrict__ out = packed + (pk_off ? pk_off[p] : (unsigned long long)p * pk_stride);
        const int rw = (pr.referenceSize + 15) >> 4, qw = (pr.querySize + 15) >> 4;
        for (int w = lane; w < rw + qw; w += 32) {
            const bool isq = w >= rw;
            const uint8_t* __restrict__ src = blob + (isq ? pr.queryIdx : pr.referenceIdx);
            const int len = isq ? pr.querySize : pr.referenceSize;
            const int k0 = (isq ? w - rw : w) << 4;
            uint32_t v = 0; bool bad = false;
            #pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k0 + k < len) { const uint8_t c = code[src[k0 + k]]; bad |= (c == 0xFF); v |= (uint32_t)(c & 3u) << (2 * k); }
            out[w] = v;
            if (bad && unknown_symbol) *unknown_symbol = 1;
        }
    }
}

// ---- batches uploaded from the host-side 2-bit sidecar (host_pack.h): the packed words arrive as they are; what the other
// kernels expect beside them is rebuilt here.  Byte layout of such a batch is synthetic: packed word w covers bytes
// [16 w, 16 w + 16), so a pair whose words start at w has referenceIdx = 16 w and queryIdx = 16 (w + ceil(R / 16)).
__global__ void __launch_bounds__(256) sidecar_expand_kernel(int n, int uni_r, int uni_q, unsigned long long stride,
                                                              const uint32_t* __restrict__ sizes, int small_sizes, const uint32_t* __restrict__ woff,
                                                              dpx_seq_pair* __restrict__ pairs, unsigned long long* __restrict__ pk_off,
                                                              unsigned long long* __restrict__ str_len) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int R = uni_r, Q = uni_q; unsigned long long w = (unsigned long long)k * stride;
    if (sizes) {
        if (small_sizes) { const uint32_t v = sizes[k]; R = (int)(v & 0xffffu); Q = (int)(v >> 16); }
        else { R = (int)sizes[2 * k]; Q = (int)sizes[2 * k + 1]; }
        w = (unsigned long long)(woff[k] - woff[0]);
    }
    dpx_seq_pair p;
    p.referenceIdx = (int32_t)(16ull * w); p.referenceSize = R;
    p.queryIdx = (int32_t)(16ull * (w + (unsigned long long)((R + 15) >> 4))); p.querySize = Q;
    pairs[k] = p;
    if (pk_off) pk_off[k] = w;
    if (str_len) str_len[k] = 3ull * ((unsigned long long)R + (unsigned long long)Q + 1ull);
}

// packed words -> bytes in the synthetic layout (only when a kernel needs the raw bytes: alignment strings, byte-compare fall-back)
__global__ void __launch_bounds__(256) unpack2_kernel(const uint32_t* __restrict__ packed, unsigned long long n_words, uint4* __restrict__ blob16, uint32_t inv4) {
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t v = packed[w];
        uint32_t o[4];
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t c = v >> (8 * q);
            const uint32_t sel = (c & 3u) | ((c >> 2 & 3u) << 4) | ((c >> 4 & 3u) << 8) | ((c >> 6 & 3u) << 12);
            o[q] = __byte_perm(inv4, 0u, sel);
        }
        blob16[w] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace dpx
