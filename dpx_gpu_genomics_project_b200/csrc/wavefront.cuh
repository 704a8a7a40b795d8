// wavefront.cuh — general warp-per-pair anti-diagonal wavefront fill (int32), all four algorithms.
//
// One warp aligns one pair.  The (Q+1)x(R+1) matrix is cut into stripes of 32*K query rows; lane t
// owns K consecutive rows of the stripe and sweeps the reference columns with a skew of t
// (column = jstart + step - t), so at every step the 32 lanes sit on one anti-diagonal of K-row
// blocks.  H (and Gotoh's vertical-gap matrix D) cross from lane t-1 to lane t by __shfl_up, from the
// last lane of a stripe to lane 0 of the next through a per-warp boundary row in global memory
// (L1/L2 resident).  Left neighbours and Gotoh's horizontal-gap matrix I never leave registers.
// Directions are packed CB bits per cell (2, Gotoh 4), K cells per lane-step, SPW steps per word and
// stored [stripe][step/SPW][lane]: one warp store = one 128-byte line (WfGeom in common.cuh).
//
// Recurrences and tie-breaks restate the reference bit for bit:
//   LNW  c++/LinearNeedlemanWunsch.cpp:89-135   (LEFT > UP > DIAG via two __vibmax_s32)
//   ANW  c++/AffineNeedlemanWunsch.cpp:167-240  (D/I tie -> GAP_OPEN; H: LEFT > UP > DIAG)
//   LSW  c++/LinearSmithWaterman.cpp:70-114     (ReLU; UP > LEFT > DIAG by equality with H)
//        end cell = first strict max in row-major order, :145-157
//   BSW  repaired semantics: LSW restricted to |i-j| <= band, everything else 0.
//   ABSW affine banded SW (not in the reference; include/dpxalign.h): ANW's D / I recurrences (tie -> GAP_OPEN) under LSW's
//        ReLU, UP > LEFT > DIAG and end-cell rules, on the band; out-of-band cells have H = 0, D = I = -inf.
#pragma once
#include "common.cuh"

namespace dpx {

struct WfArgs {
    const uint8_t* blob;                 // raw parseInput blob (bytes compared for equality, as the reference does)
    const dpx_seq_pair* pairs;
    const int32_t* order;                // schedule: order[pos] = pair id (may be null = identity)
    int first, count;                    // schedule positions [first, first+count) of this launch
    int match, mismatch, go, ge, band;   // linear algorithms: go is the gap weight; band < 0 = unbanded
    int32_t* scores;                     // [n_pairs] by pair id
    int32_t* end_rc;                     // [2*n_pairs] by pair id (may be null)
    uint32_t* tb;                        // traceback words of this launch (null when !TB)
    unsigned long long tb_stride;        // words per schedule position of this launch (geometry of (Qmax, Rmax))
    int32_t* boundary;                   // per-warp-slot boundary rows
    long long boundary_stride;           // int32 per warp slot (>= 2*(Rmax+1))
    int rmax_p1;                         // Rmax + 1 (offset of the D row inside a slot)
    unsigned int* counter;               // dynamic work counter (zeroed before launch)
};

template <int ALGO, bool TB, int K>
__global__ void __launch_bounds__(128) wf_fill_kernel(const WfArgs a) {
    constexpr bool IS_NW  = (ALGO == DPX_ALGO_LNW || ALGO == DPX_ALGO_ANW);
    constexpr bool IS_ANW = (ALGO == DPX_ALGO_ANW);
    constexpr bool IS_ASW = (ALGO == DPX_ALGO_ABSW);
    constexpr bool AFFINE = IS_ANW || IS_ASW;              // D travels down the lanes, I along the row
    constexpr bool IS_SW  = (ALGO == DPX_ALGO_LSW || ALGO == DPX_ALGO_BSW || IS_ASW);
    constexpr bool BANDED = (ALGO == DPX_ALGO_BSW || IS_ASW);
    constexpr int  CB  = AFFINE ? 4 : 2;
    constexpr int  SPW = 32 / (K * CB);
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31;
    const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int32_t* __restrict__ bH = a.boundary + slot * a.boundary_stride;
    int32_t* __restrict__ bD = bH + a.rmax_p1;
    const int g = a.go, goe = a.go + a.ge, ge = a.ge;
    const int band = BANDED ? a.band : -1;

    for (;;) {
        int pos = 0;
        if (lane == 0) pos = (int)atomicAdd(a.counter, 1u);
        pos = __shfl_sync(FULL, pos, 0);
        if (pos >= a.count) break;
        const int pid = a.order ? a.order[a.first + pos] : (a.first + pos);
        const dpx_seq_pair pr = a.pairs[pid];
        const int R = pr.referenceSize, Q = pr.querySize;
        const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
        const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
        const WfGeom geo = WfGeom::make(K, CB, Q, R, band);
        uint32_t* __restrict__ tbp = TB ? (a.tb + (unsigned long long)pos * a.tb_stride) : nullptr;

        // ---- row 0 of the matrix into the boundary row (init_matrix of each aligner) -------------
        __syncwarp();
        for (int j = lane; j <= R; j += 32) {
            int h0 = 0;
            if (ALGO == DPX_ALGO_LNW) h0 = j * g;                       // c++/LinearNeedlemanWunsch.cpp:38-41
            if (IS_ANW) h0 = (j == 0) ? 0 : a.go + j * ge;              // c++/AffineNeedlemanWunsch.cpp:50-53
            bH[j] = h0;
            if (AFFINE) bD[j] = NEG_INF;                                // D[0][j] never wins: row 1 always opens (:185-189)
        }
        __syncwarp();

        int laneBest = 0, laneRow = 0, laneCol = 0;       // SW: best of this lane's rows, row-major first
        int Hreg[K];
        int Ireg[K];
        #pragma unroll
        for (int r = 0; r < K; ++r) { Hreg[r] = 0; Ireg[r] = NEG_INF; }

        for (int s = 0; s < geo.nstripes; ++s) {
            const int i_first = s * geo.rows_per_stripe + lane * K + 1;     // first matrix row of this lane
            const int jstart = geo.jstart(s), jend = geo.jend(s);
            if (jend < jstart) continue;                                     // banded: stripe entirely right of the matrix
            const int nsteps = jend - jstart + 1 + 31;

            uint8_t qc[K];
            #pragma unroll
            for (int r = 0; r < K; ++r) qc[r] = (i_first + r <= Q) ? qry[i_first + r - 1] : (uint8_t)0;

            // left border (column jstart-1) of this lane's rows, and the cell above-left of its first row
            int diag0;
            #pragma unroll
            for (int r = 0; r < K; ++r) {
                const int i = i_first + r;
                Hreg[r] = (ALGO == DPX_ALGO_LNW) ? i * g : IS_ANW ? a.go + i * ge : 0;   // LNW :31-34, ANW :43-46, SW 0
                Ireg[r] = NEG_INF;                                                        // I[i][0] never wins (:201-205)
            }
            {
                const int i = i_first - 1;
                diag0 = (ALGO == DPX_ALGO_LNW) ? i * g : IS_ANW ? (i == 0 ? 0 : a.go + i * ge) : 0;
                // banded: the window of a later stripe starts at column jstart > 1, and the cell above-left of its first cell,
                // (s*rows, s*rows - band), sits exactly on the band's edge: it is a real cell of the previous stripe's last row
                if (BANDED && lane == 0 && s > 0 && jstart > 1) diag0 = bH[jstart - 1];
            }
            int bestH[K], bestJ[K];
            #pragma unroll
            for (int r = 0; r < K; ++r) { bestH[r] = 0; bestJ[r] = 0; }

            int botH = 0, botD = NEG_INF;
            uint32_t acc = 0;

            for (int step = 0; step < nsteps; ++step) {
                const int j = jstart + step - lane;
                int topH = __shfl_up_sync(FULL, botH, 1);
                int topD = AFFINE ? __shfl_up_sync(FULL, botD, 1) : 0;
                if (j >= jstart && j <= jend) {
                    if (lane == 0) { topH = bH[j]; if (AFFINE) topD = bD[j]; }
                    const uint8_t rc = ref[j - 1];
                    int up = topH, upD = topD, diag = diag0;
                    diag0 = topH;
                    uint32_t codes = 0;
                    #pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const int left = Hreg[r];
                        const int ds = diag + (qc[r] == rc ? a.match : a.mismatch);
                        int h; uint32_t code;
                        if (ALGO == DPX_ALGO_LNW) {
                            bool p1, p2;
                            const int m = __vibmax_s32(up + g, ds, &p1);          // up >= diag  -> QUERY_DELETION
                            h = __vibmax_s32(left + g, m, &p2);                    // left >= max -> QUERY_INSERTION
                            code = p2 ? C_LEFT : (p1 ? C_UP : C_DIAG);
                        } else if (IS_ANW) {
                            bool pd, pi, p1, p2;
                            const int dv = __vibmax_s32(up + goe, upD + ge, &pd);  // tie -> GAP_OPEN
                            const int iv = __vibmax_s32(left + goe, Ireg[r] + ge, &pi);
                            const int m = __vibmax_s32(dv, ds, &p1);
                            h = __vibmax_s32(iv, m, &p2);
                            code = (p2 ? C_LEFT : (p1 ? C_UP : C_DIAG)) | (pd ? C_DOPEN : 0u) | (pi ? C_IOPEN : 0u);
                            Ireg[r] = iv; upD = dv;
                        } else if (IS_ASW) {
                            bool pd, pi;
                            int dv = __vibmax_s32(up + goe, upD + ge, &pd);        // tie -> GAP_OPEN
                            int iv = __vibmax_s32(left + goe, Ireg[r] + ge, &pi);
                            h = __vimax3_s32_relu(dv, iv, ds);                     // max(0, D, I, diag)
                            code = (h == 0) ? C_STOP : (dv == h ? C_UP : (iv == h ? C_LEFT : C_DIAG));
                            code |= (pd ? C_DOPEN : 0u) | (pi ? C_IOPEN : 0u);
                            const unsigned d = (unsigned)(j - (i_first + r) + band);
                            if (d > (unsigned)(2 * band)) { h = 0; code = C_STOP; dv = NEG_INF; iv = NEG_INF; }   // no gap starts or continues outside the band
                            Ireg[r] = iv; upD = dv;
                            if (h > bestH[r]) { bestH[r] = h; bestJ[r] = j; }
                        } else {
                            const int ug = up + g, lg = left + g;
                            h = __vimax3_s32_relu(ug, lg, ds);                     // max(0, up, left, diag)
                            code = (h == 0) ? C_STOP : (ug == h ? C_UP : (lg == h ? C_LEFT : C_DIAG));
                            if (BANDED) {
                                const unsigned d = (unsigned)(j - (i_first + r) + band);
                                if (d > (unsigned)(2 * band)) { h = 0; code = C_STOP; }
                            }
                            if (h > bestH[r]) { bestH[r] = h; bestJ[r] = j; }      // first column of the row's max
                        }
                        diag = left; Hreg[r] = h; up = h;
                        if (TB) codes |= code << (r * CB);
                    }
                    botH = up; botD = upD;
                    if (lane == 31) { bH[j] = botH; if (AFFINE) bD[j] = botD; }
                    if (TB) {
                        const int sub = step % SPW;
                        acc |= codes << (sub * K * CB);
                        if (sub == SPW - 1 || j == jend) {
                            tbp[((size_t)s * geo.ngroups + (size_t)(step / SPW)) * 32 + lane] = acc;
                            acc = 0;
                        }
                    }
                }
            }
            if (IS_SW) {
                #pragma unroll
                for (int r = 0; r < K; ++r) {
                    const int i = i_first + r;
                    if (i <= Q && bestH[r] > laneBest) { laneBest = bestH[r]; laneRow = i; laneCol = bestJ[r]; }
                }
            }
            __syncwarp();     // boundary row written by lane 31 is read by lane 0 in the next stripe
        }

        // ---- result ------------------------------------------------------------------------------
        if (IS_NW) {
            int score;
            if (Q == 0 || R == 0) {      // pure border cell: H[Q][0] or H[0][R]
                const int n = Q + R;
                score = (ALGO == DPX_ALGO_LNW) ? n * g : (n == 0 ? 0 : a.go + n * ge);
                if (lane == 0) a.scores[pid] = score;
            } else {
                const int rowInStripe = (Q - 1) % geo.rows_per_stripe;
                if (lane == rowInStripe / K) {
                    score = 0;
                    #pragma unroll
                    for (int r = 0; r < K; ++r) if (r == rowInStripe % K) score = Hreg[r];
                    a.scores[pid] = score;
                }
            }
            if (a.end_rc && lane == 0) { a.end_rc[2 * pid] = Q; a.end_rc[2 * pid + 1] = R; }
        } else {
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int os = __shfl_xor_sync(FULL, laneBest, off);
                const int orow = __shfl_xor_sync(FULL, laneRow, off);
                const int ocol = __shfl_xor_sync(FULL, laneCol, off);
                if (os > laneBest || (os == laneBest && orow < laneRow)) { laneBest = os; laneRow = orow; laneCol = ocol; }
            }
            if (lane == 0) {
                a.scores[pid] = laneBest;
                if (a.end_rc) { a.end_rc[2 * pid] = laneRow; a.end_rc[2 * pid + 1] = laneCol; }
            }
        }
    }
}

}  // namespace dpx
