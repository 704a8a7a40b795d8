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
// Traceback, densely packed: a lane's 2M regular slots of one super-step are one field of 4M bits (2 bits per slot, order
// E[0..M-1], O[0..M-1], first slot in the highest bits), SPW = 32 / 4M steps per 32-bit word (band 64: four 8-bit fields),
// stored [step / SPW][lane]: one warp store = one 128-byte line, 0.25 B per in-band cell with no padding bits.  The 129th
// diagonal of bands with W == 32M (the extra slot of lane 31) has its own stream after the main one: 16 steps per word.
// Backtrack: one thread per pair, private shared-memory windows of the traceback refilled warp-synchronously (below).
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
    DPX_HD int field_bits() const { return 4 * M; }                      // 2 bits x 2M regular slots per lane and step
    DPX_HD int spw_shift() const { return M == 1 ? 3 : M == 2 ? 2 : 1; } // log2(steps per word): 8 / 4 / 2
    DPX_HD int spw() const { return 1 << spw_shift(); }
    DPX_HD int nsteps(int Q, int R) const { const int m = Q < R ? Q : R; return m > 0 ? ((m + W + spw() - 1) & ~(spw() - 1)) : 0; }   // whole words
    DPX_HD unsigned long long main_words(int Q, int R) const { return (unsigned long long)(nsteps(Q, R) >> spw_shift()) * 32ull; }
    DPX_HD unsigned long long words(int Q, int R) const {                // + the extra diagonal's stream, 16 steps per word, padded to a 128-byte line
        return main_words(Q, R) + (extra ? (unsigned long long)((((nsteps(Q, R) + 15) >> 4) + 31) & ~31) : 0ull);
    }
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
    constexpr int SPW = M == 1 ? 8 : M == 2 ? 4 : 2;      // super-steps per traceback word
    constexpr int FB = 4 * M;                             // bits per field
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
        const uint8_t* __restrict__ qp = qsp; const uint8_t* __restrict__ rp = rsp;
        uint32_t* __restrict__ tbw = tbp;
        uint32_t* __restrict__ tbx = (TB && EXTRA) ? (a.tb + (unsigned long long)pos * a.tb_stride + geo.main_words(Q, R)) : nullptr;   // lane 31's extra diagonal
        uint32_t accX = 0;
        const uint32_t rsh = 1u << (4 * M);

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
            const uint32_t nq1 = qp[(U) - ubase + 1], nr1 = rp[(U) - ubase + 1];   /* next step's bases (streams are padded) */ \
            qw = fma_add(qw, sixteen, nq);                                                                           \
            const uint32_t rwN = fma_add(nr1, rsh, rw >> 4);                                                         \
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
                const int m1 = __viaddmax_s32(HgE[m], (int)prmt_b32(scE, 0u, 0x8880u | (m * 0x1111u)), left);                        \
                int h = __vimax3_s32(m1, (int)fma_add((uint32_t)HgO[m], one, 1u), Z);                                \
                if (PARTIAL) h = min(h, capE[m]);                                                                    \
                DPX_BAND_TAIL(h, HgE[m], keyE[m], cE, ACC)                                                           \
            }                                                                                                        \
            if (EXTRA) {                                                                                             \
                const int m1 = __viaddmax_s32(HgX, (int)prmt_b32(scE, 0u, 0x8880u | (M * 0x1111u)), oldOlast);                       \
                const int h = __vimax3_s32(m1, zerog + 1, Z);                                                        \
                uint32_t keyX;                                                                                       \
                DPX_BAND_TAIL(h, HgX, keyX, cE, accX)                                                                \
                bestX = max(bestX, keyX);                                                                            \
            }                                                                                                        \
            int Urecv = __shfl_down_sync(FULL, HgE[0], 1);                                                           \
            if (lane == 31) Urecv = EXTRA ? HgX : zerog;                                                             \
            _Pragma("unroll")                                                                                        \
            for (int m = 0; m < M; ++m) {                                                                            \
                const int upv = (m == M - 1) ? Urecv : HgE[m + 1];                                                   \
                const int m1 = __viaddmax_s32(HgO[m], (int)prmt_b32(scO, 0u, 0x8880u | (m * 0x1111u)), HgE[m]);                      \
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
            for (int u = blk0; u < u_end; u += SPW) {
                const int ubase = u;                              // qp / rp walk the streams: constant offsets inside the step group
                uint32_t word = 0;
                #pragma unroll
                for (int q = 0; q < SPW; ++q) {
                    uint32_t acc;
                    DPX_BAND_STEP(u + q, acc)
                    if (TB) word |= acc << (q * FB);
                }
                qp += SPW; rp += SPW;
                if (TB) {
                    __stcs(tbw, word); tbw += 32;
                    // the extra diagonal: 16 steps per word, first step in the highest bits (a last partial word keeps its codes in the low bits)
                    if (EXTRA && (((u + SPW) & 15) == 0 || u + SPW >= nsteps)) { if (lane == 31) tbx[u >> 4] = accX; accX = 0; }
                }
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
    const uint8_t* blob_lo;              // first byte of the blob allocation (4-byte aligned): lower clamp of the word loads
    const dpx_seq_pair* pairs;
    const int32_t* order;
    int first, count;
    int W;
    int wshift;                          // log2(walkers per warp): 3, 4 or 5
    const int32_t* scores; const int32_t* end_rc;
    const uint32_t* tb; unsigned long long tb_stride;
    char* strings; const unsigned long long* str_off; int32_t* str_start;
};

constexpr int BAND_BT_MOVES = 32;                         // moves per round (a move lowers the super-step by <= 1)
constexpr int BAND_BT_RING_MAX = 64;                      // rows of the per-walker ring: a power of two > rows requested per round, 2 * (32 / SPW + 2):
DPX_HD int band_bt_ring(int M) { return M >= 3 ? 64 : 32; }   //   36 for bands 65..96 (2 steps per row), 20 / 12 below
constexpr int BAND_BT_LANES = 8;                          // words per row: two aligned groups of four owner lanes, placed so that the walk starts >= 2 lanes
                                                          // from either edge (the main diagonal of a band 32M sits exactly on a group boundary)
constexpr int BAND_BT_CRING = 32;                         // words of the per-walker sequence-byte rings (128 bytes each, filled 16 bytes at a time)
constexpr int BAND_BT_OUT_STRIDE = 3 * BAND_BT_MOVES + 4; // bytes per walker of the output staging (odd word stride: no bank conflicts)
// wpw = walkers per warp (8, 16 or 32): the walk is a chain of dependent shared-memory reads, so a few thousand walkers finish
// sooner spread over MORE, emptier warps (more independent instruction streams); very large batches fill every lane.
DPX_HD int band_bt_win_words(int M, int wpw) { return band_bt_ring(M) * BAND_BT_LANES * wpw; }
DPX_HD int band_bt_smem_words(int M, int wpw) { return band_bt_win_words(M, wpw) + 2 * BAND_BT_CRING * wpw + wpw * BAND_BT_OUT_STRIDE / 4; }
DPX_HD int band_bt_wpw_shift(int count, int sms) { (void)count; (void)sms; return 5; }   // measured (config 4, 10k walkers): 8 per warp 3.36 ms, 32 per warp 3.06 ms
static_assert(BAND_BT_MOVES == 32, "the flush writes one character per lane");

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// One THREAD walks one pair, one warp per block.  Every walker keeps private shared-memory rings of the traceback rows
// around its position (the aligned group of four owner lanes = 16 diagonals) and of the sequence bytes it is about to emit, filled by 16-byte cp.async one
// round AHEAD: a round requests the rows the next round can reach and waits only for what the previous round requested, so
// the HBM latency of the scattered 4-byte reads is off the walk's critical path.  Then every walker takes up to
// BAND_BT_MOVES branch-free steps out of shared memory, and the warp flushes all walkers' characters with coalesced stores.
// A walk that drifts out of its four-lane window re-centres it (one exposed round trip) or reads the word from global memory.
__global__ void __launch_bounds__(32) band_bt_kernel(const BandBtArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint32_t bt_smem[];
    const int lane = threadIdx.x;
    const int wshift = a.wshift, wpw = 1 << wshift;
    const int t = blockIdx.x * wpw + lane;
    const bool live = lane < wpw && t < a.count;
    const int W = a.W;
    const BandGeom geo = BandGeom::make(W);
    const int M = geo.M;
    const int rsh = geo.spw_shift(), FB = geo.field_bits(), spw_mask = geo.spw() - 1;
    const int need_rows = (BAND_BT_MOVES >> rsh) + 2;            // traceback rows (SPW super-steps each) one round can touch
    const int ahead_rows = 2 * need_rows;                        // rows requested per round: this round's and the next one's
    const int ring_mask = band_bt_ring(geo.M) - 1, win_words = band_bt_win_words(geo.M, wpw);
    const int wl_ = lane & (wpw - 1);                            // idle lanes (>= wpw) alias a live walker's slots but never touch them
    uint32_t* win = bt_smem + wl_ * BAND_BT_LANES;               // [row & ring_mask][walker][8 words]: two 16-byte cp.async per row
    uint32_t* rring = bt_smem + win_words + wl_ * 4;             // [(16-byte block address) & 7][walker][4 words]
    uint32_t* qring = rring + BAND_BT_CRING * wpw;
    char* outw = reinterpret_cast<char*>(bt_smem + win_words + 2 * BAND_BT_CRING * wpw);
    char* out = outw + wl_ * BAND_BT_OUT_STRIDE;                 // this walker's 3 x MOVES characters, filled from the back
    const uintptr_t blob_lo = (uintptr_t)a.blob_lo;
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
    const uint32_t rring_s = (uint32_t)__cvta_generic_to_shared(rring), qring_s = (uint32_t)__cvta_generic_to_shared(qring);
    const uint32_t out_s = (uint32_t)__cvta_generic_to_shared(out);

    int pid = 0, i = 0, j = 0, done = 1;
    const uint8_t *ref = nullptr, *qry = nullptr;
    const uint32_t* tb = nullptr;
    char *o0 = nullptr, *o1 = nullptr, *o2 = nullptr;
    long long p = 0;
    if (live) {
        pid = a.order ? a.order[a.first + t] : (a.first + t);
        const dpx_seq_pair pr = a.pairs[pid];
        ref = a.blob + pr.referenceIdx; qry = a.blob + pr.queryIdx;
        tb = a.tb + (unsigned long long)t * a.tb_stride;
        const size_t F = (size_t)pr.querySize + pr.referenceSize + 1;
        o0 = a.strings + a.str_off[pid]; o1 = o0 + F; o2 = o1 + F;
        p = (long long)F - 1;
        o0[p] = 0; o1[p] = 0; o2[p] = 0;
        if (a.scores[pid] > 0) { i = a.end_rc[2 * pid]; j = a.end_rc[2 * pid + 1]; done = 0; }
    }
    // walk state: diagonal c = j - i + W and super-step u = i + (c >> 1) - 1 move incrementally:
    //   DIAG: u-1;   UP: c+1, u-1 if c is even;   LEFT: c-1, u-1 if c is even
    int c = j - i + W, u = i + (c >> 1) - 1;
    int nsteps_w = 0;
    const uint32_t* tbx = nullptr;                               // the extra diagonal's stream of this pair
    if (live) { const dpx_seq_pair pr = a.pairs[pid]; nsteps_w = geo.nsteps(pr.querySize, pr.referenceSize); tbx = tb + geo.main_words(pr.querySize, pr.referenceSize); }
    int lane_lo = -100, row_loaded = 0;                          // rows [row_loaded, ...) of lanes lane_lo..lane_lo+2 are requested
    uintptr_t r_loaded = 0, q_loaded = 0;                        // word addresses [x_loaded, ...) are requested

    while (__any_sync(FULL, !done)) {
        if (!done) {
            bool fresh = false;
            const int row_hi = u >> rsh;
            const int owner = min((c >> 1) / M, 31);                             // (on the extra diagonal: keep lane 31's group warm)
            if ((unsigned)(owner - lane_lo - 1) >= (unsigned)(BAND_BT_LANES - 2)) {   // first round, or the walk reached an edge lane of the window
                const int lo = min(max((owner - 2) & ~3, 0), 32 - BAND_BT_LANES);
                if (lo != lane_lo) { lane_lo = lo; row_loaded = row_hi + 1; fresh = true; }
            }
            const int row_to = max(row_hi - (ahead_rows - 1), 0);
            for (int row = row_loaded - 1; row >= row_to; --row) {
                uint32_t* dst = win + (row & ring_mask) * (wpw * BAND_BT_LANES);
                cp_async16(dst, tb + (size_t)row * 32 + lane_lo);
                cp_async16(dst + 4, tb + (size_t)row * 32 + lane_lo + 4);
            }
            row_loaded = min(row_loaded, row_to);
            // sequence bytes [j - 2*MOVES, j) and [i - 2*MOVES, i) as aligned words, never below the start of the blob allocation
            const uintptr_t r_hi = ((uintptr_t)(ref + j - 1)) & ~(uintptr_t)15, q_hi = ((uintptr_t)(qry + i - 1)) & ~(uintptr_t)15;
            const uintptr_t r_to = max(((uintptr_t)(ref + j) - 2 * BAND_BT_MOVES) & ~(uintptr_t)15, blob_lo);
            const uintptr_t q_to = max(((uintptr_t)(qry + i) - 2 * BAND_BT_MOVES) & ~(uintptr_t)15, blob_lo);
            if (r_loaded == 0) { r_loaded = r_hi + 16; q_loaded = q_hi + 16; fresh = true; }
            for (uintptr_t x = r_loaded; x > r_to; ) { x -= 16; cp_async16(rring + ((x >> 4) & (BAND_BT_CRING / 4 - 1)) * (4 * wpw), (const void*)x); }
            for (uintptr_t x = q_loaded; x > q_to; ) { x -= 16; cp_async16(qring + ((x >> 4) & (BAND_BT_CRING / 4 - 1)) * (4 * wpw), (const void*)x); }
            r_loaded = min(r_loaded, r_to); q_loaded = min(q_loaded, q_to);
            cp_async_commit();
            if (fresh) cp_async_wait<0>(); else cp_async_wait<1>();
        }
        __syncwarp();
        // the walk proper: 32-bit shared-space addresses and 32-bit sequence positions keep the dependent chain short
        //   lut -> window word -> code -> next (c, u);  the two sequence bytes and the three stores hang off the side
        int cnt = 0;
        uint32_t rpos = (uint32_t)(uintptr_t)(ref + j - 1), qpos = (uint32_t)(uintptr_t)(qry + i - 1);   // low address bits index the rings
        // The traceback word of (row, owner lane) serves SPW consecutive diagonal moves: it is kept in a register and re-read from the
        // window only when the walk changes row or owner (key_cached); owner and bit position come from arithmetic on c, not a table.
        uint32_t w_cached = 0, key_cached = 0xffffffffu;
        #pragma unroll 1
        for (int mv = 0; mv < BAND_BT_MOVES; ++mv) {
            if (done) continue;
            const uint32_t k = (uint32_t)c >> 1, isO = (uint32_t)c & 1u;
            const uint32_t owner = (M == 2) ? (k >> 1) : (M == 1) ? k : k / 3u;
            const uint32_t slot = isO * (uint32_t)M + (k - owner * (uint32_t)M);
            uint32_t code;
            if (owner > 31u) {                                                   // the extra diagonal (j - i == W): its own stream, 16 steps per word
                const int nleft = min(16, nsteps_w - (u & ~15));
                code = (__ldg(tbx + (u >> 4)) >> (2 * (nleft - 1 - (u & 15)))) & 3u;
            } else {
                const uint32_t row = (uint32_t)u >> rsh, key = (row << 5) | owner;
                if (key != key_cached) {
                    const uint32_t wl = owner - (uint32_t)lane_lo;
                    w_cached = (wl < (uint32_t)BAND_BT_LANES) ? lds_u32(win_s + ((row & (uint32_t)ring_mask) << (5 + wshift)) + 4u * wl)
                                                              : __ldg(tb + (size_t)row * 32 + owner);
                    key_cached = key;
                }
                code = (w_cached >> (((uint32_t)u & (uint32_t)spw_mask) * (uint32_t)FB + 2u * (2u * (uint32_t)M - 1u - slot))) & 3u;
            }
            if (code == BD_STOP) { done = 1; continue; }
            const uint32_t rc = lds_u8(rring_s + ((rpos & 0x70u) << wshift) + (rpos & 15u));
            const uint32_t qc = lds_u8(qring_s + ((qpos & 0x70u) << wshift) + (qpos & 15u));
            ++cnt;
            const uint32_t o = out_s + (uint32_t)(BAND_BT_MOVES - cnt);
            sts_u8(o, (code == BD_UP) ? (uint32_t)'_' : rc);
            sts_u8(o + BAND_BT_MOVES, (code == BD_DIAG) ? ((rc == qc) ? (uint32_t)'*' : (uint32_t)'|') : (uint32_t)' ');
            sts_u8(o + 2 * BAND_BT_MOVES, (code == BD_LEFT) ? (uint32_t)'_' : qc);
            const int di = (code != BD_LEFT), dj = (code != BD_UP);
            u -= ((code == BD_DIAG) | ((c & 1) ^ 1));
            c += (code == BD_UP) - (code == BD_LEFT);
            i -= di; j -= dj; qpos -= (uint32_t)di; rpos -= (uint32_t)dj;
            if (i == 0 || j == 0 || (unsigned)c > (unsigned)(2 * W)) done = 1;   // the next cell is a border or out-of-band cell: H == 0
        }
        __syncwarp();
        // flush: the warp writes every walker's new characters with coalesced byte stores (positions [p - cnt, p) of each field)
        p -= cnt;
        #pragma unroll 4
        for (int w = 0; w < wpw; ++w) {
            const int n = __shfl_sync(FULL, cnt, w);
            if (n == 0) continue;
            const char* src = outw + w * BAND_BT_OUT_STRIDE + (BAND_BT_MOVES - n);
            char* d0 = (char*)__shfl_sync(FULL, (unsigned long long)(o0 + p), w);
            const long long Fw = __shfl_sync(FULL, (long long)(o1 - o0), w);
            if (lane < n) { d0[lane] = src[lane]; d0[Fw + lane] = src[BAND_BT_MOVES + lane]; d0[2 * Fw + lane] = src[2 * BAND_BT_MOVES + lane]; }
        }
        __syncwarp();
    }
    if (live) a.str_start[pid] = (int32_t)p;
}

}  // namespace dpx
