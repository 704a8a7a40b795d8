// band.cuh — BandedSmithWaterman with the band mapped onto one warp: fill (+ 2-bit traceback, end cell) and backtrack.
//
// Semantics (repaired reference; c++/BandedSmithWaterman.cpp is not executable): LinearSmithWaterman
// (c++/LinearSmithWaterman.cpp:70-114) restricted to cells with |i-j| <= W, every other cell 0; ReLU; tie-break
// UP > LEFT > DIAG; end cell = first strict maximum in row-major order (:145-157); walk stops at H == 0 (:222).
//
// Mapping.  The 2W+1 diagonals c = j - i + W are split by parity: E[k] (c = 2k, k = 0..W) and O[k] (c = 2k+1, k = 0..W-1).
// Lane t owns E[tM..tM+M-1] and O[tM..tM+M-1] with M = ceil(W/32); when W == 32M the last diagonal E[32M] is an extra
// slot of lane 31.  One super-step u computes one cell of every diagonal: first the E cells (anti-diagonal a), then the
// O cells (anti-diagonal a+1):
//       E[k] at step u is cell (i, j) = (u+1-k, u+1+k-W);  up = O[k] and left = O[k-1] of step u-1, diag = E[k] of step u-1
//       O[k] at step u is cell (u+1-k, u+2+k-W);           up = E[k+1] and left = E[k] of step u,   diag = O[k] of step u-1
// so all neighbours are in the lane's own registers except O[k-1] of the previous lane (one __shfl_up per step) and
// E[k+1] of the next lane (one __shfl_down per step).  The band never leaves the warp: no masks, no idle lanes, no
// shared memory; cells outside the matrix see never-matching pad symbols and stay 0 (top/left) or strictly below the
// maximum (bottom/right, needs gap < 0 and mismatch < 0).
//
// Arithmetic (int32): values are stored as 4*X + code, registers hold Hg = 4*(H + gap) + 1:
//       m  = viaddmax(Hg[diag], tab, Hg[left])      tab = 4*(s - gap) - 1 (int8): diag code 0, left code 1
//       h' = vimax3(m, Hg[up] + 1, 3)               up code 2, "zero" code 3: UP > LEFT > DIAG, STOP iff H == 0
//       Hg = (h' | 3) + 4*gap - 2
// the two low bits of h' ARE the traceback code.  Scores come four at a time from one PRMT over nibble windows of the
// two sequences (q + (3 - r) == 3 iff match).  End cell: key = (h'|3) * 2^kb + 2*(stepcode) + isE, one VIMNMX3.U32 per
// (E, O) slot pair keeps the highest score, then the earliest step (smallest row), then E before O (smaller column);
// folded into (score, row, col) every 2^(kb-1) steps, lanes merged at the end with the reference's row-major rule.
//
// Traceback: one 16-bit field per lane and super-step (2 bits per slot, order E[0..M-1], extra, O[0..M-1], first slot
// in the highest bits), two steps per 32-bit word, stored [step/2][lane]: one warp store = one 128-byte line.
// Backtrack: one warp per pair; the warp pulls 64 super-steps of traceback (4 KB, coalesced) and the matching sequence
// bytes into shared memory, lane 0 walks inside that window, the warp flushes the emitted characters coalesced.
#pragma once
#include "common.cuh"
#include "shortread.cuh"

namespace dpx {

constexpr uint32_t BD_DIAG = 0, BD_LEFT = 1, BD_UP = 2, BD_STOP = 3;

struct BandGeom {
    int W, M, extra;         // band, diagonals pairs per lane, 1 when lane 31 owns E[32M]
    DPX_HD static BandGeom make(int W) {
        BandGeom g; g.W = W; g.M = W <= 32 ? 1 : (W + 31) / 32; g.extra = (W == 32 * g.M) ? 1 : 0; return g;
    }
    DPX_HD int slots() const { return 2 * M + extra; }
    DPX_HD int nsteps(int Q, int R) const { const int m = Q < R ? Q : R; return m > 0 ? ((m + W + 1) & ~1) : 0; }   // even
    DPX_HD unsigned long long words(int Q, int R) const { return (unsigned long long)(nsteps(Q, R) / 2) * 32ull; }
    DPX_HD int offq() const { return 31 * M; }
    DPX_HD int offr() const { return W; }
    DPX_HD int qs_len(int Qm, int Rm) const { return (nsteps(Qm, Rm) + 31 * M + 2 + 15) & ~15; }
    DPX_HD int rs_len(int Qm, int Rm) const { return (nsteps(Qm, Rm) + 32 * M + 2 + 15) & ~15; }
};

// Padded one-byte-per-base streams: qs[n] = code(q[n - 31M]) or 4, rs[n] = 3 - code(r[n - W]) or 4.
__global__ void __launch_bounds__(256) band_prep_kernel(const uint32_t* __restrict__ packed, const unsigned long long* __restrict__ pk_off,
                                                         unsigned long long pk_stride, const dpx_seq_pair* __restrict__ pairs, int n_pairs,
                                                         int offq, int offr, int qs_len, int rs_len, uint8_t* __restrict__ qs, uint8_t* __restrict__ rs) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < n_pairs; p += nwarps) {
        const dpx_seq_pair pr = pairs[p];
        const uint32_t* __restrict__ ref = packed + (pk_off ? pk_off[p] : (unsigned long long)p * pk_stride);
        const uint32_t* __restrict__ qry = ref + ((pr.referenceSize + 15) >> 4);
        uint8_t* __restrict__ q = qs + (size_t)p * qs_len;
        uint8_t* __restrict__ r = rs + (size_t)p * rs_len;
        for (int n = lane; n < qs_len; n += 32) { const int s = n - offq; q[n] = (s >= 0 && s < pr.querySize) ? (uint8_t)get2(qry, s) : (uint8_t)4; }
        for (int n = lane; n < rs_len; n += 32) { const int s = n - offr; r[n] = (s >= 0 && s < pr.referenceSize) ? (uint8_t)(3u - get2(ref, s)) : (uint8_t)4; }
    }
}

struct BandArgs {
    const dpx_seq_pair* pairs;
    const int32_t* order;                // schedule (nullable = identity)
    int first, count;                    // schedule positions of this launch
    const uint8_t* qs; const uint8_t* rs; int qs_len, rs_len;     // padded streams, indexed by pair id
    int W;
    uint32_t lut_lo, lut_hi;             // prmt table: byte 3 = 4*(match-gap)-1, every other byte = 4*(mismatch-gap)-1
    int gadd, zerog;                     // 4*gap - 2, 4*gap + 1
    int kb; uint32_t kmul;               // position bits of the tracking keys, 2^kb
    uint32_t one, four, sixteen, minus1; // run-time multipliers (IMAD on the FMA pipe)
    int32_t* scores; int32_t* end_rc;
    uint32_t* tb; unsigned long long tb_stride;   // words per schedule position of this launch
    unsigned int* counter;
};

template <int M, bool EXTRA, bool TB>
__global__ void __launch_bounds__(128) band_sw_kernel(const BandArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool PARTIAL = !EXTRA;
    constexpr int Z = 3;
    const int lane = threadIdx.x & 31;
    const int W = a.W;
    const int gadd = a.gadd, zerog = a.zerog, kb = a.kb;
    const uint32_t kmul = a.kmul, lut_lo = a.lut_lo, lut_hi = a.lut_hi, four = a.four, sixteen = a.sixteen, one = a.one;
    const int SB = 1 << (kb - 1);
    const uint32_t minus1 = a.minus1, three_k = 3u * kmul;
    const BandGeom geo = BandGeom::make(W);

    for (;;) {
        int pos = 0;
        if (lane == 0) pos = (int)atomicAdd(a.counter, 1u);
        pos = __shfl_sync(FULL, pos, 0);
        if (pos >= a.count) break;
        const int pid = a.order ? a.order[a.first + pos] : (a.first + pos);
        const dpx_seq_pair pr = a.pairs[pid];
        const int R = pr.referenceSize, Q = pr.querySize;
        const int nsteps = geo.nsteps(Q, R);
        const uint8_t* __restrict__ qsp = a.qs + (size_t)pid * a.qs_len + (geo.offq() - lane * M);          // [u] = q base entering at step u
        const uint8_t* __restrict__ rsp = a.rs + (size_t)pid * a.rs_len + (lane * M + M - W + geo.offr());  // [u] = r base entering at step u
        uint32_t* __restrict__ tbp = TB ? (a.tb + (unsigned long long)pos * a.tb_stride + lane) : nullptr;

        int capE[M], capO[M];
        #pragma unroll
        for (int m = 0; m < M; ++m) {
            capE[m] = (lane * M + m <= W) ? 0x7fffffff : Z;
            capO[m] = (lane * M + m < W) ? 0x7fffffff : Z;
        }
        int HgE[M], HgO[M], HgX = zerog;
        uint32_t bestP[M], bestX = 0;
        #pragma unroll
        for (int m = 0; m < M; ++m) { HgE[m] = zerog; HgO[m] = zerog; bestP[m] = 0; }
        int bH = 0, bI = 0, bJ = 0;                          // lane best (score, row, col)

        uint32_t qw = 0x44444444u, rw = 0;
        #pragma unroll
        for (int n = 1; n <= M; ++n) rw |= (uint32_t)rsp[n - 1 - M] << (4 * n);
        uint32_t nq = 0, nr = 0;
        if (nsteps > 0) { nq = qsp[0]; nr = rsp[0]; }
        rw = (rw >> 4) | (nr << (4 * M));                    // window of step 0

// After h' is known: Hg = clean(h') + 4g - 2, key = clean(h') * 2^kb + code, and the 2-bit direction joins the step's field.
// With traceback the low bits are needed anyway (one LOP3), so the cleaning h' - low + 3 runs on the FMA pipe; without,
// clean(h') = h' | 3 is the cheaper single LOP3.
#define DPX_BAND_TAIL(H, HG, KEY, CODE, ACC)                                                                         \
        if (TB) {                                                                                                    \
            const uint32_t low = (uint32_t)(H) & 3u;                                                                 \
            const uint32_t tq = fma_add(low, minus1, (uint32_t)(H));              /* h' - low = 4H */                \
            HG = (int)fma_add(tq, one, (uint32_t)(gadd + 3));                                                        \
            KEY = fma_add(tq, kmul, (CODE) + three_k);                                                               \
            ACC = fma_add(ACC, four, low);                                                                           \
        } else {                                                                                                     \
            const uint32_t hc = (uint32_t)(H) | 3u;                                                                  \
            HG = (int)fma_add(hc, one, (uint32_t)gadd);                                                              \
            KEY = fma_add(hc, kmul, (CODE));                                                                         \
        }

#define DPX_BAND_STEP(U, ACC)                                                                                        \
        {                                                                                                            \
            const uint32_t nq1 = qsp[(U) + 1], nr1 = rsp[(U) + 1];             /* next step's bases (streams are padded) */ \
            qw = fma_mul(qw, sixteen) + nq;                                                                          \
            const uint32_t rwN = (rw >> 4) | (nr1 << (4 * M));                                                       \
            const uint32_t scE = prmt_b32(lut_lo, lut_hi, qw + rw), scO = prmt_b32(lut_lo, lut_hi, qw + rwN);        \
            rw = rwN; nq = nq1;                                                                                      \
            const uint32_t cE = (uint32_t)(2 * (SB - 1 - ((U) - blk0)) + 1);                                         \
            int Lrecv = __shfl_up_sync(FULL, HgO[M - 1], 1);                                                         \
            if (lane == 0) Lrecv = zerog;                                                                            \
            uint32_t keyE[M];                                                                                        \
            ACC = 0;                                                                                                 \
            const int oldOlast = HgO[M - 1];                                                                         \
            _Pragma("unroll")                                                                                        \
            for (int m = 0; m < M; ++m) {                                                                            \
                const int left = (m == 0) ? Lrecv : HgO[m - 1];                                                      \
                const int m1 = __viaddmax_s32(HgE[m], (int)(int8_t)(scE >> (8 * m)), left);                        \
                int h = __vimax3_s32(m1, (int)fma_add((uint32_t)HgO[m], one, 1u), Z);                                \
                if (PARTIAL) h = min(h, capE[m]);                                                                    \
                DPX_BAND_TAIL(h, HgE[m], keyE[m], cE, ACC)                                                           \
            }                                                                                                        \
            if (EXTRA) {                                                                                             \
                const int m1 = __viaddmax_s32(HgX, (int)(int8_t)(scE >> (8 * M)), oldOlast);                       \
                const int h = __vimax3_s32(m1, zerog + 1, Z);                                                        \
                uint32_t keyX;                                                                                       \
                DPX_BAND_TAIL(h, HgX, keyX, cE, ACC)                                                                 \
                bestX = max(bestX, keyX);                                                                            \
            }                                                                                                        \
            int Urecv = __shfl_down_sync(FULL, HgE[0], 1);                                                           \
            if (lane == 31) Urecv = EXTRA ? HgX : zerog;                                                             \
            _Pragma("unroll")                                                                                        \
            for (int m = 0; m < M; ++m) {                                                                            \
                const int upv = (m == M - 1) ? Urecv : HgE[m + 1];                                                   \
                const int m1 = __viaddmax_s32(HgO[m], (int)(int8_t)(scO >> (8 * m)), HgE[m]);                      \
                int h = __vimax3_s32(m1, (int)fma_add((uint32_t)upv, one, 1u), Z);                                   \
                if (PARTIAL) h = min(h, capO[m]);                                                                    \
                uint32_t keyO;                                                                                       \
                DPX_BAND_TAIL(h, HgO[m], keyO, cE - 1u, ACC)                                                         \
                bestP[m] = __vimax3_u32(bestP[m], keyE[m], keyO);                                                    \
            }                                                                                                        \
        }

        for (int blk0 = 0; blk0 < nsteps; blk0 += SB) {
            const int u_end = min(blk0 + SB, nsteps);
            #pragma unroll 1
            for (int u = blk0; u < u_end; u += 2) {
                uint32_t acc0, acc1;
                DPX_BAND_STEP(u, acc0)
                DPX_BAND_STEP(u + 1, acc1)
                if (TB) __stcs(tbp + (size_t)(u >> 1) * 32, acc0 | (acc1 << 16));
            }
            // fold the block's keys into the lane's (score, row, col)
            auto fold = [&](uint32_t key, int m) {
                const int H = (int)(key >> (kb + 2));
                if (H > 0) {
                    const uint32_t code = key & (kmul - 1u);
                    const int isE = (int)(code & 1u);
                    const int us = blk0 + SB - 1 - (int)(code >> 1);
                    const int k = lane * M + m;
                    const int i = us + 1 - k, j = i + 2 * k - W + (isE ? 0 : 1);
                    if (H > bH || (H == bH && (i < bI || (i == bI && j < bJ)))) { bH = H; bI = i; bJ = j; }
                }
            };
            #pragma unroll
            for (int m = 0; m < M; ++m) { fold(bestP[m], m); bestP[m] = 0; }
            if (EXTRA) { if (lane == 31) fold(bestX, M); bestX = 0; }
        }
#undef DPX_BAND_STEP
#undef DPX_BAND_TAIL

        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int oH = __shfl_xor_sync(FULL, bH, off), oI = __shfl_xor_sync(FULL, bI, off), oJ = __shfl_xor_sync(FULL, bJ, off);
            if (oH > bH || (oH == bH && oH > 0 && (oI < bI || (oI == bI && oJ < bJ)))) { bH = oH; bI = oI; bJ = oJ; }
        }
        if (lane == 0) {
            a.scores[pid] = bH;
            if (a.end_rc) { a.end_rc[2 * pid] = bH > 0 ? bI : 0; a.end_rc[2 * pid + 1] = bH > 0 ? bJ : 0; }
        }
    }
}

// ---- backtrack: one warp per pair --------------------------------------------------------------------------------
struct BandBtArgs {
    const uint8_t* blob;
    const dpx_seq_pair* pairs;
    const int32_t* order;
    int first, count;
    int W;
    const int32_t* scores; const int32_t* end_rc;
    const uint32_t* tb; unsigned long long tb_stride;
    char* strings; const unsigned long long* str_off; int32_t* str_start;
};

constexpr int BAND_BT_ROWS = 32;          // traceback rows (2 super-steps each) per window
constexpr int BAND_BT_MOVES = 64;         // moves per window (a move lowers the super-step by at most one)

__global__ void __launch_bounds__(128) band_bt_kernel(const BandBtArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_win[4][BAND_BT_ROWS * 32];
    __shared__ uint8_t s_q[4][BAND_BT_MOVES], s_r[4][BAND_BT_MOVES];
    __shared__ char s_out[4][3][BAND_BT_MOVES];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= a.count) return;
    const int pid = a.order ? a.order[a.first + t] : (a.first + t);
    const dpx_seq_pair pr = a.pairs[pid];
    const int R = pr.referenceSize, Q = pr.querySize, W = a.W;
    const BandGeom geo = BandGeom::make(W);
    const int M = geo.M, S = geo.slots();
    const int kinv = M == 1 ? 256 : M == 2 ? 128 : 86;
    const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
    const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
    const uint32_t* __restrict__ tb = a.tb + (unsigned long long)t * a.tb_stride;
    uint32_t* win = s_win[wib];

    const size_t F = (size_t)Q + R + 1;
    char* __restrict__ o0 = a.strings + a.str_off[pid];
    char* __restrict__ o1 = o0 + F;
    char* __restrict__ o2 = o1 + F;
    long long p = (long long)F - 1;
    if (lane == 0) { o0[p] = 0; o1[p] = 0; o2[p] = 0; }
    int i = 0, j = 0, done = 1;
    if (a.scores[pid] > 0) { i = a.end_rc[2 * pid]; j = a.end_rc[2 * pid + 1]; done = 0; }

    while (!done) {
        // window: traceback rows [row_lo, row_hi] and the sequence bytes the next BAND_BT_MOVES moves can touch
        const int c0 = j - i + W, k0 = c0 >> 1;
        const int row_hi = (i + k0 - 1) >> 1, row_lo = max(0, row_hi - (BAND_BT_ROWS - 1));
        __syncwarp();
        for (int rr = 0; rr <= row_hi - row_lo; ++rr) win[rr * 32 + lane] = __ldcs(tb + (size_t)(row_lo + rr) * 32 + lane);
        for (int x = lane; x < BAND_BT_MOVES; x += 32) {
            s_q[wib][x] = (i - 1 - x >= 0) ? qry[i - 1 - x] : 0;
            s_r[wib][x] = (j - 1 - x >= 0) ? ref[j - 1 - x] : 0;
        }
        __syncwarp();
        int cnt = 0;
        if (lane == 0) {
            const int i_start = i, j_start = j;
            while (cnt < BAND_BT_MOVES) {
                const int c = j - i + W;
                uint32_t code = BD_STOP;
                if (i >= 1 && j >= 1 && c >= 0 && c <= 2 * W) {
                    const int k = c >> 1, isO = c & 1;
                    const int u = i + k - 1;
                    if ((u >> 1) < row_lo) break;                         // next window
                    const int owner = min((k * kinv) >> 8, 31);                 // k / M for M <= 3
                    const int slot = isO ? (M + geo.extra + (k - owner * M)) : (k - owner * M);    // processing order E.., extra, O..
                    const uint32_t w = win[((u >> 1) - row_lo) * 32 + owner];
                    code = (w >> ((u & 1) * 16 + 2 * (S - 1 - slot))) & 3u;
                }
                if (code == BD_STOP) { done = 1; break; }
                const char rc = (char)s_r[wib][j_start - j], qc = (char)s_q[wib][i_start - i];
                if (code == BD_DIAG) { s_out[wib][0][cnt] = rc; s_out[wib][1][cnt] = (rc == qc) ? '*' : '|'; s_out[wib][2][cnt] = qc; --i; --j; }
                else if (code == BD_UP) { s_out[wib][0][cnt] = '_'; s_out[wib][1][cnt] = ' '; s_out[wib][2][cnt] = qc; --i; }
                else { s_out[wib][0][cnt] = rc; s_out[wib][1][cnt] = ' '; s_out[wib][2][cnt] = '_'; --j; }
                ++cnt;
                if (i == 0 || j == 0) { done = 1; break; }                // the next cell is a border cell: H == 0
            }
        }
        cnt = __shfl_sync(FULL, cnt, 0); done = __shfl_sync(FULL, done, 0);
        i = __shfl_sync(FULL, i, 0); j = __shfl_sync(FULL, j, 0);
        __syncwarp();
        for (int x = lane; x < cnt; x += 32) {               // x-th emitted character sits at position p - 1 - x
            o0[p - 1 - x] = s_out[wib][0][x]; o1[p - 1 - x] = s_out[wib][1][x]; o2[p - 1 - x] = s_out[wib][2][x];
        }
        p -= cnt;
    }
    if (lane == 0) a.str_start[pid] = (int32_t)p;
}

}  // namespace dpx
