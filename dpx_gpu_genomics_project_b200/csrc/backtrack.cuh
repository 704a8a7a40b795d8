// backtrack.cuh — GPU traceback: walks the packed direction slab written by the fill kernels and emits
// the three alignment strings (REF / REL / QRY) exactly as the reference prints them.
//
// Walk rules restate the reference:
//   LNW  c++/LinearNeedlemanWunsch.cpp:137-223 (= backtrackNW, c++/backtrack.cpp:21-81): from (Q,R)
//        while (i != 0 || j != 0); row 0 => QUERY_INSERTION, column 0 => QUERY_DELETION (init_matrix :31-41).
//   ANW  c++/AffineNeedlemanWunsch.cpp:242-403 (= backtrackANW, c++/backtrack.cpp:214-356): 3-state walk
//        while (i != 0 && j != 0), then pad rows as deletions, then columns as insertions (:366-378).
//   LSW  c++/LinearSmithWaterman.cpp:163-226 (= backtrackSW, c++/backtrack.cpp:83-144): from the first-max
//        cell, apply the cell's direction, stop when the next cell's H is 0; score 0 => three empty strings.
// Characters: '*' match, '|' mismatch, ' ' gap in REL, '_' gap symbol.
//
// One thread walks one pair (the walk is a chain of dependent 4-byte loads from an L2-resident slab, so
// parallelism comes from walking many pairs at once).  Strings are written right-to-left into the pair's
// slot: three fields of F = Q+R+1 bytes, each ending in NUL; str_start[pid] = offset of the first
// character inside each field.
#pragma once
#include "common.cuh"

namespace dpx {

struct BtArgs {
    const uint8_t* blob;
    const dpx_seq_pair* pairs;
    const int32_t* order;
    int first, count;
    int K, band;                         // geometry of the slab that was written (band < 0 = unbanded)
    const int32_t* scores;
    const int32_t* end_rc;               // SW start cells
    const uint32_t* tb;
    unsigned long long tb_stride;        // words per schedule position (as in WfArgs)
    char* strings;                       // output slab
    const unsigned long long* str_off;   // [n_pairs] byte offset of the pair's slot
    int32_t* str_start;                  // [n_pairs] out
};

template <int CB>
__device__ __forceinline__ uint32_t wf_code(const uint32_t* __restrict__ tb, const WfGeom& g, int i, int j) {
    const int ii = i - 1;
    const int s = ii / g.rows_per_stripe;
    const int rem = ii - s * g.rows_per_stripe;
    const int lane = rem / g.K, r = rem - lane * g.K;
    const int js = g.jstart(s);
    if (j < js || j > g.jend(s)) return C_STOP;           // banded: outside the stripe's column window => H == 0
    const int step = j - js + lane;
    const int grp = step / g.SPW, sub = step - grp * g.SPW;
    const uint32_t w = __ldg(tb + ((size_t)s * g.ngroups + grp) * 32 + lane);
    return (w >> ((sub * g.K + r) * CB)) & ((1u << CB) - 1u);
}

template <int ALGO>
__global__ void __launch_bounds__(128) bt_walk_kernel(const BtArgs a) {
    constexpr int CB = (ALGO == DPX_ALGO_ANW || ALGO == DPX_ALGO_ABSW) ? 4 : 2;
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.count) return;
    const int pid = a.order ? a.order[a.first + pos] : (a.first + pos);
    const dpx_seq_pair pr = a.pairs[pid];
    const int R = pr.referenceSize, Q = pr.querySize;
    const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
    const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
    const WfGeom geo = WfGeom::make(a.K, CB, Q, R, (ALGO == DPX_ALGO_BSW || ALGO == DPX_ALGO_ABSW) ? a.band : -1);
    const uint32_t* __restrict__ tb = a.tb + (unsigned long long)pos * a.tb_stride;

    const size_t F = (size_t)Q + R + 1;
    char* __restrict__ o0 = a.strings + a.str_off[pid];
    char* __restrict__ o1 = o0 + F;
    char* __restrict__ o2 = o1 + F;
    long long p = (long long)F - 1;
    o0[p] = 0; o1[p] = 0; o2[p] = 0;

    auto emit_diag = [&](int i, int j) { const char rc = ref[j - 1], qc = qry[i - 1]; --p; o0[p] = rc; o1[p] = (rc == qc) ? '*' : '|'; o2[p] = qc; };
    auto emit_up   = [&](int i)        { --p; o0[p] = '_'; o1[p] = ' '; o2[p] = qry[i - 1]; };
    auto emit_left = [&](int j)        { --p; o0[p] = ref[j - 1]; o1[p] = ' '; o2[p] = '_'; };

    if (ALGO == DPX_ALGO_LNW) {
        int i = Q, j = R;
        while (i != 0 || j != 0) {
            const uint32_t c = (i == 0) ? C_LEFT : (j == 0) ? C_UP : wf_code<CB>(tb, geo, i, j);
            if (c == C_DIAG) { emit_diag(i, j); --i; --j; }
            else if (c == C_UP) { emit_up(i); --i; }
            else { emit_left(j); --j; }
        }
    } else if (ALGO == DPX_ALGO_ANW) {
        int i = Q, j = R, state = 0;     // 0 SCORING, 1 INSERTION, 2 DELETION
        while (i != 0 && j != 0) {
            const uint32_t c = wf_code<CB>(tb, geo, i, j);
            if (state == 0) {
                const uint32_t d = c & 3u;
                if (d == C_DIAG) { emit_diag(i, j); --i; --j; }
                else if (d == C_UP) state = 2;
                else state = 1;
            } else if (state == 1) {
                state = (c & C_IOPEN) ? 0 : 1;
                emit_left(j); --j;
            } else {
                state = (c & C_DOPEN) ? 0 : 2;
                emit_up(i); --i;
            }
        }
        while (i > 0) { emit_up(i); --i; }
        while (j > 0) { emit_left(j); --j; }
    } else if (ALGO == DPX_ALGO_ABSW) {
        // affine banded SW: ANW's three states, LSW's stop rule (STOP code <=> H == 0, checked on arrival in SCORING)
        if (a.scores[pid] > 0) {
            int i = a.end_rc[2 * pid], j = a.end_rc[2 * pid + 1], state = 0;
            while (i != 0 && j != 0) {
                const uint32_t c = wf_code<CB>(tb, geo, i, j);
                if (state == 0) {
                    const uint32_t d = c & 3u;
                    if (d == C_STOP) break;
                    if (d == C_DIAG) { emit_diag(i, j); --i; --j; }
                    else if (d == C_UP) state = 2;
                    else state = 1;
                } else if (state == 1) {
                    state = (c & C_IOPEN) ? 0 : 1;
                    emit_left(j); --j;
                } else {
                    state = (c & C_DOPEN) ? 0 : 2;
                    emit_up(i); --i;
                }
            }
        }
    } else {
        if (a.scores[pid] > 0) {
            int i = a.end_rc[2 * pid], j = a.end_rc[2 * pid + 1];
            uint32_t c = wf_code<CB>(tb, geo, i, j);
            while (c != C_STOP) {
                if (c == C_DIAG) { emit_diag(i, j); --i; --j; }
                else if (c == C_UP) { emit_up(i); --i; }
                else { emit_left(j); --j; }
                if (i == 0 || j == 0) break;
                c = wf_code<CB>(tb, geo, i, j);
            }
        }
    }
    a.str_start[pid] = (int32_t)p;
}

// string_offsets[3*i + k] = slot offset + k * F + first character  (the ABI's view of the slab)
__global__ void __launch_bounds__(256) str_offsets_kernel(const dpx_seq_pair* __restrict__ pairs, int n, const unsigned long long* __restrict__ str_off,
                                                          const int32_t* __restrict__ str_start, unsigned long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long F = (unsigned long long)pairs[i].querySize + (unsigned long long)pairs[i].referenceSize + 1ull;
    const unsigned long long b = str_off[i] + (unsigned long long)str_start[i];
    out[3 * i] = b; out[3 * i + 1] = b + F; out[3 * i + 2] = b + 2 * F;
}

// Compacted string output: pair i owns 3 * (len_i + 1) bytes (REF, REL, QRY, each NUL-terminated), pairs in index order.
__global__ void __launch_bounds__(256) str_len_kernel(const dpx_seq_pair* __restrict__ pairs, int n, const int32_t* __restrict__ str_start,
                                                      unsigned long long* __restrict__ len3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { len3[i] = 0; return; }
    const unsigned long long F = (unsigned long long)pairs[i].querySize + (unsigned long long)pairs[i].referenceSize + 1ull;
    len3[i] = 3ull * (F - (unsigned long long)str_start[i]);          // F - start = alignment length + NUL
}

__global__ void __launch_bounds__(256) str_compact_kernel(const dpx_seq_pair* __restrict__ pairs, int n, const char* __restrict__ slab,
                                                          const unsigned long long* __restrict__ str_off, const int32_t* __restrict__ str_start,
                                                          const unsigned long long* __restrict__ coff, char* __restrict__ out,
                                                          unsigned long long* __restrict__ offs, unsigned long long base = 0ull) {   // base: added to the offsets only
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        const unsigned long long F = (unsigned long long)pairs[i].querySize + (unsigned long long)pairs[i].referenceSize + 1ull;
        const unsigned long long L = F - (unsigned long long)str_start[i];
        const char* __restrict__ src = slab + str_off[i] + (unsigned long long)str_start[i];
        char* __restrict__ dst = out + coff[i];
        for (int k = 0; k < 3; ++k)
            for (unsigned long long x = lane; x < L; x += 32) dst[k * L + x] = src[k * F + x];
        if (lane < 3) offs[3 * i + lane] = base + coff[i] + lane * L;
    }
}

// ---- formatted output: the reference's stdout block of every pair, written on the device --------------------------------
//   "<index> | <score>\n"  then (with strings)  REF "\n" REL "\n" QRY "\n"     c++/LinearNeedlemanWunsch.cpp:207-213;
//   a Smith-Waterman score of 0 prints three empty lines (c++/LinearSmithWaterman.cpp:253-257) = three empty strings here.
__device__ __forceinline__ int dec_digits(long long v) {           // characters of printf("%lld")
    int n = v < 0 ? 2 : 1;
    unsigned long long a = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
    while (a >= 10) { a /= 10; ++n; }
    return n;
}
__device__ __forceinline__ char* put_dec(char* p, long long v) {   // writes the digits, returns the end
    const int n = dec_digits(v);
    unsigned long long a = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
    for (int k = n - 1; k >= (v < 0 ? 1 : 0); --k) { p[k] = (char)('0' + a % 10); a /= 10; }
    if (v < 0) p[0] = '-';
    return p + n;
}

__global__ void __launch_bounds__(256) text_len_kernel(const dpx_seq_pair* __restrict__ pairs, int n, long long first_index,
                                                       const int32_t* __restrict__ scores, const int32_t* __restrict__ str_start /* null: no strings */,
                                                       unsigned long long* __restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { len[i] = 0; return; }
    unsigned long long l = (unsigned long long)(dec_digits(first_index + i) + 3 + dec_digits(scores[i]) + 1);
    if (str_start) {
        const unsigned long long F = (unsigned long long)pairs[i].querySize + (unsigned long long)pairs[i].referenceSize + 1ull;
        l += 3ull * (F - (unsigned long long)str_start[i]);           // alignment length + newline, three times
    }
    len[i] = l;
}

__global__ void __launch_bounds__(256) text_write_kernel(const dpx_seq_pair* __restrict__ pairs, int n, long long first_index,
                                                         const int32_t* __restrict__ scores, const char* __restrict__ slab,
                                                         const unsigned long long* __restrict__ str_off, const int32_t* __restrict__ str_start,
                                                         const unsigned long long* __restrict__ toff, char* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        char* __restrict__ dst = out + toff[i];
        const int hdr = dec_digits(first_index + i) + 3 + dec_digits(scores[i]) + 1;
        if (lane == 0) {
            char* p = put_dec(dst, first_index + i);
            p[0] = ' '; p[1] = '|'; p[2] = ' ';
            p = put_dec(p + 3, scores[i]);
            p[0] = '\n';
        }
        if (str_start) {
            const unsigned long long F = (unsigned long long)pairs[i].querySize + (unsigned long long)pairs[i].referenceSize + 1ull;
            const unsigned long long L = F - 1ull - (unsigned long long)str_start[i];
            const char* __restrict__ src = slab + str_off[i] + (unsigned long long)str_start[i];
            for (int k = 0; k < 3; ++k) {
                char* __restrict__ d = dst + hdr + k * (L + 1);
                for (unsigned long long x = lane; x < L; x += 32) d[x] = src[k * F + x];
                if (lane == 0) d[L] = '\n';
            }
        }
    }
}

// sort key of the schedule: longer queries first, then longer references (descending via bitwise not)
__global__ void __launch_bounds__(256) sched_keys_kernel(const dpx_seq_pair* __restrict__ pairs, int n, unsigned long long* __restrict__ keys, int32_t* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ~(((unsigned long long)(uint32_t)pairs[i].querySize << 32) | (unsigned long long)(uint32_t)pairs[i].referenceSize);
    ids[i] = i;
}

// in-band cell count of the banded algorithm: sum over rows of |{j in [1,R] : |i-j| <= W}|
__global__ void __launch_bounds__(256) band_cells_kernel(const dpx_seq_pair* __restrict__ pairs, int n, int W, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long c = 0;
    for (int p = warp; p < n; p += nwarps) {
        const long long Q = pairs[p].querySize, R = pairs[p].referenceSize;
        for (long long i = 1 + lane; i <= Q; i += 32) {
            const long long lo = i - W < 1 ? 1 : i - W, hi = i + W > R ? R : i + W;
            if (hi >= lo) c += (unsigned long long)(hi - lo + 1);
        }
    }
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0 && c) atomicAdd(out, c);
}

}  // namespace dpx
