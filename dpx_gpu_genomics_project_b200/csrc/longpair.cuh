// longpair.cuh — LinearSmithWaterman score + end cell for ONE very long pair (BASELINE config 5), int32.
//
// The reference cannot run this case at all (8 B/cell full matrix: c++/LinearSmithWaterman.cpp:32-45); the semantics are
// those of its recurrence (:70-114) and of its end-cell rule, first strict maximum in row-major order (:145-157).
//
// Mapping: a linear systolic array of warps.  Warp w owns a block of CW = 32*K consecutive reference COLUMNS (lane l owns
// K of them, H of the previous row in registers) and sweeps the query rows top to bottom with a lane skew of one row, so
// the 32 lanes sit on an anti-diagonal; the right-most H of a lane reaches the next lane by __shfl_up.  The right edge of
// warp w (one H per row) flows to warp w+1 through a CHANNEL: a ring of int32 in global memory plus two counters,
// `progress` (rows published by the producer, st.release) and `credit` (rows consumed by the consumer, for back-pressure).
// Every warp of a launch must be co-resident (cooperative launch).  Channels make no assumption about where their
// memory lives, so the same kernel serves
//   * one GPU, one pass               : channels are L2-resident rings between neighbouring warps;
//   * one GPU, several passes         : the last warp's channel is a full-length array that the next pass reads;
//   * several GPUs (column stripes)   : the last warp's ring and progress counter are PEER memory of the next GPU
//                                       (NVLink P2P stores, st.release.sys), its credit counter is written back by the peer.
// A watchdog turns a would-be deadlock into an error flag instead of a hung GPU.
#pragma once
#include "common.cuh"

namespace dpx {

// A channel carries one H value per query row from a producer warp to a consumer warp.  Protocol = flag-in-data (the
// same idea as NCCL's LL protocol): every ring entry is ONE aligned 8-byte word {row tag : 32 | H : 32}; an 8-byte store is
// atomic, so the consumer simply re-reads an entry until its tag equals the row it expects — no memory fence, no separate
// "progress" flag, no ld.acquire (which would invalidate L1 for the whole SM).  Back-pressure is a relaxed `credit`
// counter written by the consumer; the producer looks at it only when it is about to lap the ring.
struct LongChan {
    unsigned long long* ring; // null: constant zero input (matrix column 0) / discarded output
    long long* credit;        // rows 1..*credit have been consumed (consumer writes, relaxed); null: no back-pressure needed
    long long size;           // ring entries; row i lives at i % size (size > Q: a full-length array, never wraps)
};

struct LongArgs {
    const uint8_t* ref;       // device: reference bytes of this launch's stripe (R_local bytes)
    const uint8_t* qry;       // device: query bytes (Q)
    long long Q, R_local;
    long long col0;           // local 0-based column of warp 0's first column
    long long col_offset;     // global matrix column of local column 0 is col_offset + 1
    int match, mismatch, gap;
    int nwarps;
    const LongChan* chans;    // [nwarps + 1]: warp w reads chans[w], writes chans[w + 1]
    int32_t* best_score;      // [nwarps]
    long long* best_row;      // [nwarps]
    long long* best_col;      // [nwarps]  (global, 1-based)
    int* error_flag;
    int system_scope;         // 1: a channel crosses GPUs
    // TABLE kernels: ref / qry hold 2-bit codes (one per byte) and scores come from a per-column byte table
    int tab_match, tab_mismatch;   // match - gap, mismatch - gap (both fit int8)
    uint32_t sixteen;              // 16, passed as data so that h*16 + (15-k) stays one IMAD on the FMA pipe
    // checkpoint mode (longtrace.cuh): every warp also leaves the column it READS, H[i][left edge of its block], in
    // ck_base[(ck_first + w - 1) * ck_stride + i] — 32 rows per coalesced store, straight from the block prologue's registers
    int32_t* ck_base;              // null: no checkpoints
    long long ck_stride;
    long long ck_first;            // global index of this launch's warp 0
    // ... and (CK kernels) every lane leaves its K columns of the rows i = m * 2^rk_shift: H[i][j] at rk_base[(m - 1) * rk_stride + j - 1]
    int32_t* rk_base;
    long long rk_stride;
    int rk_shift;
};

__device__ __forceinline__ uint32_t prmt_b32l(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));   // generic mode: nibble msb = replicate sign
    return d;
}
__device__ __forceinline__ uint32_t fma_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));    // b is run-time data: stays an IMAD (FMA pipe)
    return d;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_volatile_s64(const long long* p) {
    long long v;
    asm volatile("ld.volatile.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s64(long long* p, long long v) {
    asm volatile("st.volatile.global.s64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

constexpr int LONG_WATCHDOG = 1 << 22;   // polls before a stuck channel raises the error flag

// Nothing in the row loop depends on L1: reference bases sit in registers; the query and the incoming channel move 32
// rows at a time with coalesced loads (issued one block ahead) and reach the lane that needs them by warp shuffle; lane 31
// stores its finished row straight into the outgoing ring.  With PACK the travelling H and the query base share one
// 32-bit word (H < 2^23), so a row step costs two shuffles.  Ring sizes are powers of two (index = row & (size-1)).
// TABLE: sequences are 2-bit codes; column k keeps a 4-byte table T[k] = (score - gap) for query codes 0..3 and the
// registers hold hg = H + gap of the previous row, so a cell is PRMT + 2 x VIADDMNMX on the ALU pipe + one FMA-pipe add:
//     s' = prmt(T[k], sel(q))                      t = max(hg_diag + s', hg_up)                 (VIADDMNMX)
//     h  = max(left + gap, t, 0)  (VIADDMNMX.RELU)  hg = h + gap                                 (IMAD)
// and the end cell is tracked per LANE (a max3 tree per row step, a rare branch when the lane's maximum grows).
// !TABLE: byte compare + per-column tracking, for alphabets with more than four symbols.
template <int K, bool PACK, bool TABLE, bool CK = false>
__global__ void __launch_bounds__(128) long_sw_kernel(const LongArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= a.nwarps) return;
    const LongChan cin = a.chans[w], cout = a.chans[w + 1];
    const int Q = (int)a.Q;
    const long long cfirst = a.col0 + (long long)w * (32 * K) + (long long)lane * K;    // local 0-based
    const int g = a.gap, ma = a.match, mi = a.mismatch;
    constexpr int KPB = K > 16 ? 5 : 4, KP = 1 << KPB;           // column codes of the per-lane keys: h * KP + (KP-1-k)
    const uint32_t sixteen = a.sixteen * (K > 16 ? 2u : 1u);       // KP as run-time data (IMAD on the FMA pipe)
    const unsigned imask = cin.ring ? (unsigned)(cin.size - 1) : 0u, omask = cout.ring ? (unsigned)(cout.size - 1) : 0u;
    const bool has_in = cin.ring != nullptr, has_out = cout.ring != nullptr;

    uint32_t rc[K]; bool cv[K];
    int Hc[K], bestS[TABLE ? 1 : K], bestR[TABLE ? 1 : K];   // rows fit in int32 (Q < 2^31 is checked by the host)
    #pragma unroll
    for (int k = 0; k < K; ++k) {
        cv[k] = (cfirst + k) < a.R_local;
        if (TABLE) {
            // table word: byte c = (c == ref code ? match : mismatch) - gap; columns beyond the stripe never match
            const uint32_t xs = (uint32_t)(a.tab_mismatch & 0xff), ms = (uint32_t)(a.tab_match & 0xff);
            uint32_t t = xs * 0x01010101u;
            if (cv[k]) { const uint32_t c = a.ref[cfirst + k] & 3u; t = (t & ~(0xffu << (8 * c))) | (ms << (8 * c)); }
            rc[k] = t;
            Hc[k] = g;                                                 // hg of matrix row 0: 0 + gap
        } else {
            rc[k] = cv[k] ? (uint32_t)a.ref[cfirst + k] : 0x100u;       // 0x100 never equals a query byte
            Hc[k] = 0;
        }
    }
    #pragma unroll
    for (int k = 0; k < (TABLE ? 1 : K); ++k) { bestS[k] = 0; bestR[k] = 0; }
    int bestC = 0;                                                  // TABLE: column index (0..K-1) of the lane's best cell
    int lastH = 0, leftprev = TABLE ? g : 0;
    uint32_t qc = 0;
    long long credit = 0;
    bool dead = false;
    const int nsteps = Q + 31;
    const int rk_mask = CK ? (1 << a.rk_shift) - 1 : 0;

    // software prefetch of the first block
    uint32_t qnext = (lane < Q) ? (uint32_t)__ldcg(a.qry + lane) : 0u;
    unsigned long long enext = 0;
    if (has_in && 1 + lane <= Q) enext = ld_volatile_u64(cin.ring + ((unsigned)(1 + lane) & imask));

    for (int sb = 0; sb < nsteps && !dead; sb += 32) {
        // ---- block prologue: this block's 32 query bases and 32 left-boundary values (prefetched), then prefetch the next ----
        const uint32_t qbuf = qnext;
        int rin = 0;
        if (has_in && sb < Q) {
            const int row = sb + 1 + lane;
            unsigned long long e = enext;
            int spins = 0;
            while (__any_sync(FULL, row <= Q && (unsigned)(e >> 32) != (unsigned)row)) {     // not published yet: poll
                if (row <= Q && (unsigned)(e >> 32) != (unsigned)row) e = ld_volatile_u64(cin.ring + ((unsigned)row & imask));
                if (++spins > LONG_WATCHDOG || (((spins & 1023) == 0) && *(volatile const int*)a.error_flag)) { dead = true; break; }
            }
            if (dead) break;
            rin = (int)(unsigned)e;
            if (a.ck_base != nullptr && row <= Q) a.ck_base[(a.ck_first + w - 1) * a.ck_stride + row] = rin;
            if (cin.credit && lane == 0) st_volatile_s64(cin.credit, (sb + 32 < Q) ? sb + 32 : Q);   // those rows now live in registers
        }
        {
            const int nb = sb + 32;
            qnext = (nb + lane < Q) ? (uint32_t)__ldcg(a.qry + nb + lane) : 0u;
            if (has_in && nb + 1 + lane <= Q) enext = ld_volatile_u64(cin.ring + ((unsigned)(nb + 1 + lane) & imask));
        }
        // back-pressure once per block: rows up to sb+2 leave this warp during the block
        if (has_out && cout.credit && (long long)(sb + 2) - credit >= cout.size - 64) {
            int spins = 0;
            if (lane == 0)
                while ((long long)(sb + 2) - (credit = ld_volatile_s64(cout.credit)) >= cout.size - 64) {
                    if (++spins > LONG_WATCHDOG || *(volatile const int*)a.error_flag) { dead = true; break; }
                    __nanosleep(40);
                }
            credit = __shfl_sync(FULL, credit, 0);
            dead = __any_sync(FULL, dead);
            if (dead) break;
        }
        const uint32_t pinbuf = PACK ? (((uint32_t)rin << 8) | qbuf) : 0u;
        // CK: a checkpoint row m * 2^rk_shift is crossed by the skewed lanes in the two blocks around it; those take the
        // checked loop, which carries the dump code, so that the unrolled steady loop stays as small as the score-only kernel's
        const bool ckblk = CK && ((sb & rk_mask) == 0 || ((sb + 32) & rk_mask) == 0);
        const bool steady = (sb >= 31) && (sb + 32 <= Q) && !ckblk;    // every lane is inside the matrix for all 32 steps

#define DPX_LONG_STEP(CHECKED)                                                                                   \
        {                                                                                                        \
            const int s = sb + t;                                                                                \
            const int i = s - lane + 1;                                                                          \
            int leftH;                                                                                           \
            if (PACK) {                                                                                          \
                uint32_t pk = __shfl_up_sync(FULL, ((uint32_t)lastH << 8) | qc, 1);                              \
                const uint32_t p0 = __shfl_sync(FULL, pinbuf, t);                                                \
                if (lane == 0) pk = p0;                                                                          \
                leftH = (int)(pk >> 8); qc = pk & 0xffu;                                                         \
            } else {                                                                                             \
                const uint32_t q_up = __shfl_up_sync(FULL, qc, 1), q0 = __shfl_sync(FULL, qbuf, t);              \
                qc = (lane == 0) ? q0 : q_up;                                                                    \
                leftH = __shfl_up_sync(FULL, lastH, 1);                                                          \
                const int l0 = __shfl_sync(FULL, rin, t);                                                        \
                if (lane == 0) leftH = l0;                                                                       \
            }                                                                                                    \
            if (!(CHECKED) || (i >= 1 && i <= Q)) {                                                              \
                int diag = leftprev, left = leftH;                                                               \
                if (TABLE) {                                                                                     \
                    leftprev = leftH + g;                                                                        \
                    const uint32_t sel = qc * 0x1111u + 0x8880u;        /* byte q, sign-replicated upwards */    \
                    /* keys h*16 + (15-k): the row maximum with its FIRST column falls out of a max3 tree */     \
                    int m0 = 0, m1 = 0;                                                                          \
                    _Pragma("unroll")                                                                            \
                    for (int k = 0; k < K; ++k) {                                                                \
                        const int sc = (int)prmt_b32l(rc[k], 0u, sel);                                           \
                        const int t2 = __viaddmax_s32(diag, sc, Hc[k]);                                          \
                        const int h = __viaddmax_s32_relu(left, g, t2);                                          \
                        diag = Hc[k]; Hc[k] = h + g; left = h;                                                   \
                        const int key = (int)fma_u32((uint32_t)h, sixteen, (uint32_t)(KP - 1 - k)); /* FMA pipe */ \
                        if (k & 1) m1 = __vimax3_s32(m1, m0, key); else m0 = key;                                \
                    }                                                                                            \
                    if (K & 1) m1 = max(m1, m0);                                                                 \
                    if ((m1 >> KPB) > bestS[0]) { bestS[0] = m1 >> KPB; bestR[0] = i; bestC = KP - 1 - (m1 & (KP - 1)); } \
                } else {                                                                                         \
                    leftprev = leftH;                                                                            \
                    _Pragma("unroll")                                                                            \
                    for (int k = 0; k < K; ++k) {                                                                \
                        const int up = Hc[k];                                                                    \
                        const int h = __vimax3_s32_relu(diag + (qc == rc[k] ? ma : mi), up + g, left + g);       \
                        diag = up; Hc[k] = h; left = h;                                                          \
                        if (h > bestS[k]) { bestS[k] = h; bestR[k] = i; } /* first row of the column's max */    \
                    }                                                                                            \
                }                                                                                                \
                lastH = left;                                                                                    \
                if (CK && (CHECKED) && (i & rk_mask) == 0) {        /* a checkpoint row: once per 2^rk_shift rows */ \
                    int32_t* dst = a.rk_base + ((long long)(i >> a.rk_shift) - 1) * a.rk_stride + a.col_offset + cfirst; \
                    _Pragma("unroll")                                                                            \
                    for (int k = 0; k < K; ++k) if (cv[k]) dst[k] = TABLE ? Hc[k] - g : Hc[k];                   \
                }                                                                                                \
                if (has_out && lane == 31)                                                                       \
                    st_volatile_u64(cout.ring + ((unsigned)i & omask), ((unsigned long long)(unsigned)i << 32) | (unsigned)left); \
            }                                                                                                    \
        }

        if (steady) {
            #pragma unroll 4
            for (int t = 0; t < 32; ++t) DPX_LONG_STEP(false)
        } else {
            #pragma unroll 1
            for (int t = 0; t < 32; ++t) { if (sb + t >= nsteps) break; DPX_LONG_STEP(true) }
        }
#undef DPX_LONG_STEP
    }
    if (__any_sync(FULL, dead) && lane == 0) atomicExch(a.error_flag, 1);

    // ---- best cell of this warp: higher score, then smaller row, then smaller column --------------------------
    int bs = 0; long long br = 0, bc = 0;
    if (TABLE) {
        // invalid columns never match, so they stay below the valid cell to their left and cannot be the lane's first maximum
        if (bestS[0] > 0) { bs = bestS[0]; br = bestR[0]; bc = a.col_offset + cfirst + bestC + 1; }
    } else {
        #pragma unroll
        for (int k = 0; k < K; ++k)
            if (cv[k] && (bestS[k] > bs || (bestS[k] == bs && bs > 0 && bestR[k] < br))) { bs = bestS[k]; br = bestR[k]; bc = a.col_offset + cfirst + k + 1; }
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int os = __shfl_xor_sync(FULL, bs, off);
        const long long orow = __shfl_xor_sync(FULL, br, off), ocol = __shfl_xor_sync(FULL, bc, off);
        if (os > bs || (os == bs && os > 0 && (orow < br || (orow == br && ocol < bc)))) { bs = os; br = orow; bc = ocol; }
    }
    if (lane == 0) { a.best_score[w] = bs; a.best_row[w] = br; a.best_col[w] = bc; }
}

}  // namespace dpx
