// allmax.cuh — LinearSmithWaterman in the reference's BACKTRACK_ALL mode (c++/LinearSmithWaterman.h:9):
// EVERY cell holding the maximum starts a walk (c++/LinearSmithWaterman.cpp:126-143, queued bottom-right to top-left), each walk
// follows the one direction per cell of the single-path mode (UP if up == H, else LEFT if left == H, else DIAG, :104-108) until
// H == 0 (:222), and the finished alignments come out ordered by their number of moves (the reference advances all queued walks
// one move per turn, :163-226), ties in queue order.
//
// This is the reference's teaching / inspection mode, not a throughput path: the score matrices are kept whole (int32, 4 B per
// cell, all pairs of the call back to back) and everything else reads them:
//   am_fill_kernel    warp per pair, anti-diagonal by anti-diagonal (lanes stride over the cells of a diagonal); maximum, then the
//                     number of cells that hold it
//   am_list_kernel    warp per pair: the maximum cells in descending row-major order (ballot + popc keep the order)
//   am_walk_kernel    thread per start cell: number of moves of its walk (directions re-derived from H)
//   am_emit_kernel    thread per start cell: the three lines, written from the back, at the offset the host computed from the
//                     sorted (moves, queue order) list
#pragma once
#include "common.cuh"

namespace dpx {

struct AmArgs {
    const uint8_t* blob;                 // parseInput blob (bytes compared for equality)
    const dpx_seq_pair* pairs;
    int n_pairs;
    int match, mismatch, gap;
    int32_t* H;                          // all matrices back to back; pair k at H + moff[k], (R + 1) columns per row
    const long long* moff;
    int32_t* best;                       // [n_pairs] maximum of the matrix
    int32_t* count;                      // [n_pairs] cells holding it (0 when the maximum is 0)
    const long long* soff;               // [n_pairs] first start slot of the pair
    int* start_i; int* start_j;          // [n_starts]
    int* start_pair;                     // [n_starts]
    long long n_starts;
    long long* moves;                    // [n_starts]
    const long long* toff;               // [n_starts] byte offset of the alignment's REF line in the text
    uint8_t* text;
};

__global__ void __launch_bounds__(128) am_fill_kernel(const AmArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (k >= a.n_pairs) return;
    const dpx_seq_pair pr = a.pairs[k];
    const int R = pr.referenceSize, Q = pr.querySize;
    const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
    const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
    int32_t* H = a.H + a.moff[k];
    const long long W = (long long)R + 1;
    const int g = a.gap;
    int best = 0;
    for (int d = 2; d <= Q + R; ++d) {                      // cells (i, j) with i + j == d depend on diagonals d - 1 and d - 2 only
        const int ilo = max(1, d - R), ihi = min(Q, d - 1);
        for (int i = ilo + lane; i <= ihi; i += 32) {
            const int j = d - i;
            const int up = H[(long long)(i - 1) * W + j] + g, left = H[(long long)i * W + (j - 1)] + g;
            const int diag = H[(long long)(i - 1) * W + (j - 1)] + (qry[i - 1] == ref[j - 1] ? a.match : a.mismatch);
            const int h = __vimax3_s32_relu(up, left, diag);     // max(0, up, left, diag), c++/LinearSmithWaterman.cpp:97-100
            H[(long long)i * W + j] = h;
            best = max(best, h);
        }
        __syncwarp();                                        // the diagonal is complete (and visible) before the next one reads it
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(FULL, best, off));
    int cnt = 0;
    if (best > 0) {
        const long long N = W * ((long long)Q + 1);
        for (long long x = lane; x < N; x += 32) cnt += (H[x] == best);
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(FULL, cnt, off);
    }
    if (lane == 0) { a.best[k] = best; a.count[k] = cnt; }
}

__global__ void __launch_bounds__(128) am_list_kernel(const AmArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (k >= a.n_pairs) return;
    const int best = a.best[k];
    if (best <= 0) return;
    const dpx_seq_pair pr = a.pairs[k];
    const long long W = (long long)pr.referenceSize + 1, N = W * ((long long)pr.querySize + 1);
    const int32_t* H = a.H + a.moff[k];
    long long slot = a.soff[k];
    for (long long base = N - 1; base >= 0; base -= 32) {   // descending row-major index = rows descending, columns descending (:126-127)
        const long long x = base - lane;
        const bool hit = x >= 0 && H[x] == best;
        const unsigned m = __ballot_sync(FULL, hit);
        if (hit) {
            const long long s = slot + __popc(m & ((1u << lane) - 1u));
            a.start_i[s] = (int)(x / W); a.start_j[s] = (int)(x % W); a.start_pair[s] = k;
        }
        slot += __popc(m);
    }
}

// One move of the walk from (i, j), H[i][j] > 0: the direction rule of c++/LinearSmithWaterman.cpp:104-108 read back from the scores.
__device__ __forceinline__ uint32_t am_dir(const int32_t* H, long long W, int i, int j, int g) {
    const int h = H[(long long)i * W + j];
    if (H[(long long)(i - 1) * W + j] + g == h) return C_UP;
    if (H[(long long)i * W + (j - 1)] + g == h) return C_LEFT;
    return C_DIAG;
}

__global__ void __launch_bounds__(128) am_walk_kernel(const AmArgs a) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_starts) return;
    const int k = a.start_pair[s];
    const long long W = (long long)a.pairs[k].referenceSize + 1;
    const int32_t* H = a.H + a.moff[k];
    int i = a.start_i[s], j = a.start_j[s];
    long long n = 0;
    do {
        const uint32_t d = am_dir(H, W, i, j, a.gap);
        i -= (d != C_LEFT); j -= (d != C_UP);
        ++n;
    } while (H[(long long)i * W + j] != 0);                 // stop when the next cell scores 0 (:222); row 0 and column 0 are 0
    a.moves[s] = n;
}

__global__ void __launch_bounds__(128) am_emit_kernel(const AmArgs a) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_starts) return;
    const int k = a.start_pair[s];
    const dpx_seq_pair pr = a.pairs[k];
    const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
    const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
    const long long W = (long long)pr.referenceSize + 1, L = a.moves[s];
    const int32_t* H = a.H + a.moff[k];
    uint8_t* o0 = a.text + a.toff[s]; uint8_t* o1 = o0 + L + 1; uint8_t* o2 = o1 + L + 1;
    o0[L] = o1[L] = o2[L] = '\n';
    int i = a.start_i[s], j = a.start_j[s];
    for (long long pos = L - 1; pos >= 0; --pos) {
        const uint32_t d = am_dir(H, W, i, j, a.gap);
        const uint8_t qi = qry[i - 1], rj = ref[j - 1];
        o0[pos] = d == C_UP ? (uint8_t)'_' : rj;
        o1[pos] = d == C_DIAG ? (qi == rj ? (uint8_t)'*' : (uint8_t)'|') : (uint8_t)' ';
        o2[pos] = d == C_LEFT ? (uint8_t)'_' : qi;
        i -= (d != C_LEFT); j -= (d != C_UP);
    }
}

}  // namespace dpx
