// pairwf.cuh — Needleman-Wunsch fills WITH traceback in packed int16x2: LinearNeedlemanWunsch and
// AffineNeedlemanWunsch (Gotoh), two pairs per warp, directions carried in the low bits of the scores.
//
// Reference semantics restated bit for bit:
//   LNW  c++/LinearNeedlemanWunsch.cpp:89-135   two __vibmax_s32: UP if up >= diag, then LEFT if left >= max(up, diag)
//   ANW  c++/AffineNeedlemanWunsch.cpp:167-240  D/I: tie -> GAP_OPEN (:185-213); H: UP if D >= diag, LEFT if I >= max (:229-233)
//        borders c++/AffineNeedlemanWunsch.cpp:43-53, c++/LinearNeedlemanWunsch.cpp:31-41
//
// Mapping.  One warp aligns TWO pairs at once: pair A in the low, pair B in the high int16 half of every register.
// Lane t keeps K consecutive query rows in registers and sweeps the reference columns one step behind lane t-1
// (anti-diagonal wavefront of K-row blocks); the bottom row of a lane reaches the next lane by __shfl_up (H, and D for
// Gotoh).  32*K rows are one pass; longer queries take several passes, the last lane's row handed over through a
// per-warp shared-memory row.
//
// Arithmetic: every value X is stored as 4*X + BIAS + code.  The two low bits are a tie-break code, so a plain
// max picks the winner AND records who won:
//   Gotoh   ds = Ho[diag] + tab              4*(Hdiag + s)        code 0   (IMAD, FMA pipe; tab = 4*(s - goe) - 3 from PRMT)
//           D' = viaddmax(Dc[up], 4ge, Ho[up])   extend code 1 / open code 3: tie -> OPEN; bit 1 = "D opened here"
//           I' = viaddmax(Ic[left], 4ge, Ho[left]) extend code 2 / open code 3: tie -> OPEN; bit 0 = "I opened here"
//           Dc = D' & ~2 (code 1), Ic = I' & ~1 (code 2)
//           h' = vimax3(ds, Dc, Ic)          code 0 DIAG < 1 UP < 2 LEFT: exactly the reference's LEFT > UP > DIAG ties
//           Ho = (h' | 3) + 4*goe            "H + open + extend", code 3: what the right / lower / diagonal neighbours need
//   => per cell-pair, with traceback: PRMT, 3 DPX, 4 LOP3 (clean D / I, the code, the two open flags in one) on the ALU pipe and
//      5 IMAD (ds, two to bank the 4-bit code, two for Ho = (h' - code) + 4*goe + 3); without: PRMT, 3 DPX, one LOP3 and 2 IMAD.
//   Linear  m  = viaddmax(Hg[diag], tab, Hg[up])   diag code 0 / up code 1;  h' = viaddmax(Hg[left], 1, m)  left code 2
//           Hg = (h' | 3) + 4g - 2           "H + gap", code 1
//   => per cell-pair 5 ALU-pipe + 3 FMA-pipe instructions.
//   Smith-Waterman (c++/LinearSmithWaterman.cpp:70-114, tie-break UP > LEFT > DIAG by equality with H, NONE iff H == 0):
//           m  = viaddmax(Hg[diag], tab, Hg[left])   diag code 0 / left code 1;  h' = vimax3(m, Hg[up] + 1, zero)  up code 2,
//           "zero" = BIAS + 3 is the STOP code;  Hg = (h' - code) + 4g + 1.  End cell (first strict maximum in row-major order,
//           :145-157): every register row keeps its running maximum of clean(h') with one predicate-returning VIMNMX.S16x2 and
//           the step at which it last grew (two SEL); rows, lanes and passes are merged by (score desc, row asc, column asc).
//   => per cell-pair 7 ALU-pipe + 4 FMA-pipe instructions.
// Alphabets of 5..8 symbols (the reference's data sets use '0'..'4') run the int32 instantiation with BOTH table registers of a
// row serving one pair (WIDE): PRMT indexes eight entries, so nothing else in the kernel changes.
// All adds that are not fused into a DPX instruction are plain 32-bit IMADs on the FMA pipe: every stored value is
// biased positive (>= the largest |constant|), so adding a negative constant ALWAYS carries out of the low half (the
// constant's high half is pre-decremented) and adding a table entry (>= 0) NEVER does.
//
// Traceback: CB = 4 bits per cell (Gotoh: dir | I-open << 2 | D-open << 3) or 2 (linear: dir); each lane-step banks K
// cells per pair into W = K*CB/16 words whose low / high halves belong to pair A / B (row r of the block sits in
// word r / (16/CB), nibble or bit pair (16/CB - 1 - r % (16/CB))), stored [pass][step][w][lane]: one warp store = one
// 128-byte line.  Algorithmic traceback bytes per cell: 0.5 (Gotoh), 0.25 (linear).
#pragma once
#include "common.cuh"
#include "shortread.cuh"

namespace dpx {

constexpr uint32_t PW_DIAG = 0, PW_UP = 1, PW_LEFT = 2;     // direction codes of this kernel family
constexpr uint32_t PW_IOPEN = 4, PW_DOPEN = 8;              // = 4 * (D' & I' & 3): see the Gotoh step
constexpr uint32_t PW_SW_DIAG = 0, PW_SW_LEFT = 1, PW_SW_UP = 2, PW_SW_STOP = 3;   // Smith-Waterman codes of this family

struct PwGeom {
    int K, CB, W;                // rows per lane, code bits, words per lane-step
    DPX_HD static PwGeom make(int K, int CB) { PwGeom g; g.K = K; g.CB = CB; g.W = K * CB / 16; return g; }
    DPX_HD int passes(int Qw) const { return (Qw + 32 * K - 1) / (32 * K); }
    DPX_HD int nsteps(int Rw) const { return (Rw + 31 + 1) & ~1; }          // column steps incl. pipeline drain, even
    DPX_HD unsigned long long words(int Qw, int Rw) const { return (unsigned long long)passes(Qw) * nsteps(Rw) * W * 32ull; }
};

struct PwArgs {
    const uint32_t* packed;              // 2-bit packed sequences (pack.cuh)
    const unsigned long long* pk_off;    // [n_pairs] word offsets, or null: pair p starts at p * pk_stride
    unsigned long long pk_stride;
    const dpx_seq_pair* pairs;
    const int32_t* order;                // schedule (nullable = identity)
    int first, count;                    // schedule positions [first, first + count) of this launch; slot s = positions first+2s, first+2s+1
    uint32_t lut_lo, lut_hi;             // low byte of lut_lo = mismatch entry, low byte of lut_hi = match entry (4*(s - open) - code)
    uint32_t ext2;                       // Gotoh: packed 4*ge; linear: packed 1
    uint32_t addc;                       // carry-compensated add constant: Gotoh 4*goe, linear 4*g - 2   (applied to h' | 3)
    uint32_t addc3;                      // the same + 3                                                  (applied to h' - code)
    uint32_t minus1;                     // 0xffffffff: run-time multiplier for FMA-pipe subtractions
    uint32_t zero2;                      // Smith-Waterman: packed BIAS + 3 (H = 0 with the STOP code)
    uint32_t one, four, sixteen;         // run-time multipliers: keep shifts / adds on the FMA pipe as IMAD
    int b0, b1, bstep;                   // border(idx) = idx == 0 ? b0 : b1 + bstep * idx   (stored form, one half)
    int dec_sub, dec_add;                // score = ((half - dec_sub) >> 2) + dec_add
    int32_t* scores;
    int32_t* end_rc;                     // nullable; NW: (Q, R)
    uint32_t* tb;                        // traceback words of this launch (slot-major); unused when !TB
    unsigned long long tb_stride;        // words per slot
    unsigned int* counter;
    int bnd_stride, rsel_stride;         // per-warp shared-memory strides (uint32 / uint16 entries)
    uint32_t* bnd_global;                // null: boundary rows live in shared memory
    const uint8_t* codes;                // WIDE (5..8 symbols): the blob with every byte replaced by its code 0..7
};

__device__ __forceinline__ uint32_t fma_mul(uint32_t a, uint32_t m) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(d) : "r"(a), "r"(m));
    return d;
}
// max.s16x2 with the two "a >= b" predicates (VIMNMX.S16x2 Rd, P1, P2).  Same PTX as CUDA's __vibmax_s16x2, but with
// early-clobber outputs: the header's asm reads `a` after writing the result, so an in-place update (a = vibmax(a, b))
// lets the register allocator overlap them and the predicates come out always true.
__device__ __forceinline__ uint32_t vibmax_s16x2_safe(uint32_t a, uint32_t b, bool* ge_hi, bool* ge_lo) {
    uint32_t val, phi, plo;
    asm("{.reg .pred pu, pv; \n\t"
        ".reg .s16 rs0, rs1, rs2, rs3; \n\t"
        "max.s16x2 %0, %3, %4; \n\t"
        "mov.b32 {rs0, rs1}, %0; \n\t"
        "mov.b32 {rs2, rs3}, %3; \n\t"
        "setp.eq.s16 pv, rs0, rs2; \n\t"
        "setp.eq.s16 pu, rs1, rs3; \n\t"
        "selp.b32 %1, 1, 0, pu; \n\t"
        "selp.b32 %2, 1, 0, pv;} \n\t"
        : "=&r"(val), "=&r"(phi), "=&r"(plo) : "r"(a), "r"(b));
    *ge_hi = (bool)phi; *ge_lo = (bool)plo;
    return val;
}

// One kernel source for both arithmetic widths: PACKED = two pairs per warp in int16x2 halves, !PACKED = one pair in int32
// (same recurrences, codes and traceback layout with the upper half of every word unused; for scores beyond the int16 budget).
template <bool P> __device__ __forceinline__ uint32_t pw_addmax(uint32_t a, uint32_t b, uint32_t c) {
    return P ? __viaddmax_s16x2(a, b, c) : (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);
}
template <bool P> __device__ __forceinline__ uint32_t pw_max3(uint32_t a, uint32_t b, uint32_t c) {
    return P ? __vimax3_s16x2(a, b, c) : (uint32_t)__vimax3_s32((int)a, (int)b, (int)c);
}
template <bool P> __device__ __forceinline__ uint32_t pw_bmax(uint32_t a, uint32_t b, bool* ge_hi, bool* ge_lo) {
    if (P) return vibmax_s16x2_safe(a, b, ge_hi, ge_lo);
    *ge_hi = true; *ge_lo = (int)a >= (int)b;
    return (uint32_t)max((int)a, (int)b);
}
template <bool P> __device__ __forceinline__ uint32_t pw_val(int v) { return P ? (uint32_t)(v & 0xffff) * 0x00010001u : (uint32_t)v; }

template <int ALGO, bool TB, int K, bool PACKED, bool GBND, bool WIDE>
__global__ void __launch_bounds__(128) pw_nw_kernel(const PwArgs a) {
    static_assert(!(WIDE && PACKED), "alphabets of 5..8 symbols use both table registers for ONE pair: int32 only");
    constexpr uint32_t REP1 = PACKED ? 0x00010001u : 1u, REP2 = 2u * REP1, REP3 = 3u * REP1;   // a small constant in every lane of a register
    extern __shared__ uint32_t pw_smem[];
    constexpr bool AFF = (ALGO == DPX_ALGO_ANW);
    constexpr bool SW = (ALGO == DPX_ALGO_LSW);
    constexpr int CB = AFF ? 4 : 2;
    constexpr int CPH = 16 / CB;                   // cells per half-word
    constexpr int W = K / CPH;
    static_assert(K % CPH == 0, "K must fill whole traceback words");
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // pass-to-pass boundary rows: shared memory while they leave room for >= 5 blocks per SM, else a per-warp global buffer (L2)
    // (GBND is a template flag so that the common case keeps provably-shared pointers: LDS / STS, not generic loads)
    uint32_t* __restrict__ bndH = GBND ? a.bnd_global + ((size_t)blockIdx.x * 4 + wib) * a.bnd_stride * (AFF ? 2 : 1)
                                       : pw_smem + (size_t)wib * a.bnd_stride * (AFF ? 2 : 1);
    uint32_t* __restrict__ bndD = bndH + a.bnd_stride;
    uint16_t* __restrict__ rsel = reinterpret_cast<uint16_t*>(pw_smem + (GBND ? 0 : (size_t)4 * a.bnd_stride * (AFF ? 2 : 1))) + (size_t)wib * a.rsel_stride;
    const uint32_t one = a.one, ext2 = a.ext2, addc = a.addc, addc3 = a.addc3, minus1 = a.minus1, zero2 = a.zero2;
    const uint32_t ms1 = a.lut_hi & 0xffu, xs4 = (a.lut_lo & 0xffu) * 0x01010101u;     // table entries: match, mismatch
    const uint32_t four = a.four, sixteen = a.sixteen;
    const int n_slots = PACKED ? (a.count + 1) >> 1 : a.count;     // packed: two pairs per warp; unpacked (int32): one
    const PwGeom geo = PwGeom::make(K, CB);

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = (int)atomicAdd(a.counter, 1u);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot >= n_slots) break;
        const int posA = a.first + (PACKED ? 2 * slot : slot);
        const int pa = a.order ? a.order[posA] : posA;
        int pb = -1, RB = 0, QB = 0;
        const dpx_seq_pair prA = a.pairs[pa];
        const int RA = prA.referenceSize, QA = prA.querySize;
        const uint32_t* refA = WIDE ? nullptr : a.packed + (a.pk_off ? a.pk_off[pa] : (unsigned long long)pa * a.pk_stride);
        const uint32_t* qryA = WIDE ? nullptr : refA + ((RA + 15) >> 4);
        const uint32_t *refB = refA, *qryB = qryA;
        const uint8_t* ref8 = WIDE ? a.codes + prA.referenceIdx : nullptr;      // WIDE: one code byte (0..7) per base, blob indexing
        const uint8_t* qry8 = WIDE ? a.codes + prA.queryIdx : nullptr;
        if (PACKED && 2 * slot + 1 < a.count) {
            pb = a.order ? a.order[posA + 1] : posA + 1;
            const dpx_seq_pair prB = a.pairs[pb];
            RB = prB.referenceSize; QB = prB.querySize;
            refB = a.packed + (a.pk_off ? a.pk_off[pb] : (unsigned long long)pb * a.pk_stride); qryB = refB + ((RB + 15) >> 4);
        }
        const int Rw = max(RA, RB), Qw = max(QA, QB);
        const int passes = geo.passes(Qw);
        const int nsteps2 = geo.nsteps(Rw);
        uint32_t* __restrict__ tbp = TB ? (a.tb + (unsigned long long)slot * a.tb_stride) : nullptr;

        // ---- per-warp column table: entry e <-> column j = e - 31; pads never match -----------------------------
        __syncwarp();
        for (int e = lane; e < nsteps2 + 32; e += 32) {
            const int j = e - 31;
            // selector of the per-row score tables: nibble 0 = reference base of pair A (bytes 0..3 of the row's table pair),
            // nibble 2 = 4 + base of pair B (bytes 4..7), nibbles 1 / 3 = the same index | 8 (sign extension; entries are >= 0).
            // Pad columns select the sign of byte 0 / 4 = 0: below any real entry.
            const uint32_t nA = (j >= 1 && j <= RA) ? (WIDE ? (uint32_t)ref8[j - 1] : get2(refA, j - 1)) : 8u;
            const uint32_t nB = (PACKED && j >= 1 && j <= RB) ? 4u + get2(refB, j - 1) : 12u;
            rsel[e] = PACKED ? (uint16_t)(nA | ((nA | 8u) << 4) | (nB << 8) | ((nB | 8u) << 12))
                             : (uint16_t)(nA | ((nA | 8u) * 0x1110u));          // int32: one base, its sign in the three upper bytes
        }
        __syncwarp();

        // where H[Q][R] of each pair appears: pass, lane, register row
        const int ppA = (QA > 0) ? (QA - 1) / (32 * K) : -1, glA = (QA > 0) ? ((QA - 1) / K) & 31 : -1, rrA = (QA > 0) ? (QA - 1) % K : 0;
        const int ppB = (QB > 0) ? (QB - 1) / (32 * K) : -1, glB = (QB > 0) ? ((QB - 1) / K) & 31 : -1, rrB = (QB > 0) ? (QB - 1) % K : 0;

        int swA = 0, swRowA = 0, swColA = 0, swB = 0, swRowB = 0, swColB = 0;     // SW: this lane's best (clean stored score, row, column)

        for (int p = 0; p < passes; ++p) {
            const int i0 = p * 32 * K + lane * K;                // rows above this lane's block; its rows are i0+1 .. i0+K
            uint32_t ta[K], tb[K], hA[K], hB[K], Ic[AFF ? K : 1];   // ta / tb: table entry of this row's base against bases 0..3
            uint32_t bestv[SW ? K : 1]; int stA[SW ? K : 1], stB[SW ? K : 1];        // SW: per-row running maximum and the step it last grew at
            #pragma unroll
            for (int r = 0; r < (SW ? K : 1); ++r) { bestv[r] = 0; stA[r] = 0; stB[r] = 0; }
            #pragma unroll
            for (int r = 0; r < K; ++r) {
                const int i = i0 + r;                            // 0-based query index
                const uint32_t qa = (i < QA) ? (WIDE ? (uint32_t)qry8[i] : get2(qryA, i)) : 8u;
                const uint32_t qb = (PACKED && i < QB) ? get2(qryB, i) : 8u;
                ta[r] = (qa < 4u) ? ((xs4 & ~(0xffu << (8 * qa))) | (ms1 << (8 * qa))) : xs4;
                if (WIDE) tb[r] = (qa >= 4u && qa < 8u) ? ((xs4 & ~(0xffu << (8 * (qa - 4u)))) | (ms1 << (8 * (qa - 4u)))) : xs4;   // symbols 4..7 of the SAME pair
                else      tb[r] = !PACKED ? 0u : (qb < 4u) ? ((xs4 & ~(0xffu << (8 * qb))) | (ms1 << (8 * qb))) : xs4;
                hA[r] = pw_val<PACKED>(a.b1 + a.bstep * (i + 1));       // column 0 border of matrix row i+1
                hB[r] = hA[r];
                if constexpr (AFF) Ic[r] = REP2;                    // I[i][0] never wins: column 1 always opens (:201-205)
            }
            uint32_t botH = hA[K - 1], botD = REP1;
            uint32_t topprev = pw_val<PACKED>(i0 == 0 ? a.b0 : a.b1 + a.bstep * i0);   // H[i0][0]: diagonal of (i0+1, 1)
            const bool more = (p + 1 < passes);
            uint32_t* __restrict__ tbpp = TB ? (tbp + (size_t)p * nsteps2 * W * 32 + lane) : nullptr;

#define DPX_PW_STEP(S, OLD, NEW)                                                                                     \
            {                                                                                                        \
                const int j = (S) - lane + 1;                                                                        \
                uint32_t topH = __shfl_up_sync(FULL, botH, 1);                                                       \
                uint32_t topD = AFF ? __shfl_up_sync(FULL, botD, 1) : 0u;                                            \
                uint32_t acc[W > 0 ? W : 1];                                                                         \
                _Pragma("unroll")                                                                                    \
                for (int w = 0; w < W; ++w) acc[w] = 0;                                                              \
                if (j >= 1) {                                                                                        \
                    if (lane == 0) {                                                                                 \
                        /* row 0 border (D[0][j] never wins, :185-189), or the previous pass's last row; columns beyond the   \
                           duo are pads whose row was never written: give them border values so they stay in range */      \
                        if (p == 0 || j > Rw) { topH = pw_val<PACKED>(a.b1 + a.bstep * j); topD = REP1; }                \
                        else { topH = bndH[j]; if (AFF) topD = bndD[j]; }                                            \
                    }                                                                                                \
                    const uint32_t rs = rsel[(S) - lane + 32];                                                       \
                    uint32_t up = topH, upD = topD, diag = topprev;                                                  \
                    topprev = topH;                                                                                  \
                    _Pragma("unroll")                                                                                \
                    for (int r = 0; r < K; ++r) {                                                                    \
                        const uint32_t sc = prmt_b32(ta[r], tb[r], rs);                                              \
                        uint32_t h;                                                                                  \
                        if constexpr (AFF) {                                                                         \
                            const uint32_t ds = fma_add(diag, one, sc);                                              \
                            const uint32_t Dn = pw_addmax<PACKED>(upD, ext2, up);                                     \
                            const uint32_t In = pw_addmax<PACKED>(Ic[r], ext2, OLD[r]);                               \
                            /* the codes only matter to the traceback: without it D / I keep whatever code won (the 4X part is the \
                               same either way, and (h' | 3) below renormalises what the neighbours read) */                        \
                            const uint32_t Dc = TB ? (Dn & ~REP2) : Dn, Icn = TB ? (In & ~REP1) : In;                        \
                            h = pw_max3<PACKED>(ds, Dc, Icn);                                                         \
                            upD = Dc; Ic[r] = Icn;                                                                   \
                            if (TB) {   /* nibble = dir + 4 * (I opened) + 8 * (D opened).  D' ends in 01 (extend) or 11 (open),   \
                                           I' in 10 or 11: D' & I' & 3 = 2 * (D opened) + (I opened), one LOP3 */          \
                                const uint32_t t = h & REP3;                                                  \
                                const uint32_t f = Dn & In & REP3;                                            \
                                acc[r / CPH] = fma_add(acc[r / CPH], sixteen, t);                                    \
                                acc[r / CPH] = fma_add(f, four, acc[r / CPH]);                                       \
                                NEW[r] = fma_add(fma_add(t, minus1, h), one, addc3);   /* (h' | 3) + addc without a LOP3: the ALU pipe is the full one */ \
                            }                                                                                        \
                        } else if constexpr (SW) {                                                                   \
                            const uint32_t m = pw_addmax<PACKED>(diag, sc, OLD[r]);          /* diag 0 / left 1 */    \
                            h = pw_max3<PACKED>(m, fma_add(up, one, ext2), zero2);           /* up 2 / zero 3 */      \
                            const uint32_t t = h & REP3;                                                      \
                            const uint32_t hq = fma_add(t, minus1, h);                      /* clean: 4H + BIAS */   \
                            if (TB) acc[r / CPH] = fma_add(acc[r / CPH], four, t);                                   \
                            NEW[r] = fma_add(hq, one, addc3);                                                        \
                            bool keepHi, keepLo;                                            /* best >= hq: no new maximum */ \
                            bestv[r] = pw_bmax<PACKED>(bestv[r], hq, &keepHi, &keepLo);                            \
                            stA[r] = keepLo ? stA[r] : (S);                                                          \
                            stB[r] = keepHi ? stB[r] : (S);                                                          \
                        } else {                                                                                     \
                            const uint32_t m = pw_addmax<PACKED>(diag, sc, up);                                       \
                            h = pw_addmax<PACKED>(OLD[r], ext2, m);                                                   \
                            if (TB) {                                                                                \
                                const uint32_t t = h & REP3;                                                  \
                                acc[r / CPH] = fma_add(acc[r / CPH], four, t);                                       \
                                NEW[r] = fma_add(fma_add(t, minus1, h), one, addc3);                                 \
                            }                                                                                        \
                        }                                                                                            \
                        diag = OLD[r];                                                                               \
                        if (!TB && !SW) NEW[r] = fma_add(h | REP3, one, addc);                                \
                        up = NEW[r];                                                                                 \
                    }                                                                                                \
                    botH = up; botD = upD;                                                                           \
                    if (lane == 31 && more) { bndH[j] = botH; if (AFF) bndD[j] = botD; }                             \
                }                                                                                                    \
                if (TB) {                                                                                            \
                    _Pragma("unroll")                                                                                \
                    for (int w = 0; w < W; ++w) __stcs(tbpp + ((size_t)(S) * W + w) * 32, acc[w]);                   \
                }                                                                                                    \
            }

            // H[Q][R] of a pair appears in lane glX at step RX + glX - 1 of pass ppX.  The step loop is cut after the step
            // pair that contains it, so the capture stays out of the hot loop: after a pair (s, s+1) hB holds step s, hA step s+1.
            const int sA = (p == ppA && RA > 0) ? RA + glA - 1 : -1, sB = (p == ppB && RB > 0) ? RB + glB - 1 : -1;
            const int eA = sA >= 0 ? (sA | 1) + 1 : nsteps2, eB = sB >= 0 ? (sB | 1) + 1 : nsteps2;
            int s = 0;
            #pragma unroll 1
            for (int seg = 0; seg < 3; ++seg) {
                const int s_end = seg == 0 ? min(eA, eB) : seg == 1 ? max(eA, eB) : nsteps2;
                #pragma unroll 1
                for (; s < s_end; s += 2) {
                    DPX_PW_STEP(s, hA, hB)
                    DPX_PW_STEP(s + 1, hB, hA)
                }
                if (!SW && sA >= 0 && s == eA && lane == glA) {
                    uint32_t v = 0;
                    #pragma unroll
                    for (int r = 0; r < K; ++r) if (r == rrA) v = (sA & 1) ? hA[r] : hB[r];
                    a.scores[pa] = (((PACKED ? (int)(v & 0xffffu) : (int)v) - a.dec_sub) >> 2) + a.dec_add;
                }
                if (!SW && sB >= 0 && s == eB && lane == glB) {
                    uint32_t v = 0;
                    #pragma unroll
                    for (int r = 0; r < K; ++r) if (r == rrB) v = (sB & 1) ? hA[r] : hB[r];
                    a.scores[pb] = (((int)(v >> 16) - a.dec_sub) >> 2) + a.dec_add;
                }
            }
#undef DPX_PW_STEP
            if constexpr (SW) {
                // rows of this pass into the lane's best: higher score, then smaller row (rows ascend with r and with the pass)
                #pragma unroll
                for (int r = 0; r < K; ++r) {
                    const int vA = PACKED ? (int)(bestv[r] & 0xffffu) : (int)bestv[r], vB = PACKED ? (int)(bestv[r] >> 16) : 0;
                    if (i0 + r < QA && vA > swA) { swA = vA; swRowA = i0 + r + 1; swColA = stA[r] - lane + 1; }
                    if (i0 + r < QB && vB > swB) { swB = vB; swRowB = i0 + r + 1; swColB = stB[r] - lane + 1; }
                }
            }
            __syncwarp();
        }

        if constexpr (SW) {
            // merge the lanes: higher score, then smaller row, then smaller column; score 0 (= the bias) has no end cell
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int oA = __shfl_xor_sync(FULL, swA, off), orA = __shfl_xor_sync(FULL, swRowA, off), ocA = __shfl_xor_sync(FULL, swColA, off);
                const int oB = __shfl_xor_sync(FULL, swB, off), orB = __shfl_xor_sync(FULL, swRowB, off), ocB = __shfl_xor_sync(FULL, swColB, off);
                if (oA > swA || (oA == swA && (orA < swRowA || (orA == swRowA && ocA < swColA)))) { swA = oA; swRowA = orA; swColA = ocA; }
                if (oB > swB || (oB == swB && (orB < swRowB || (orB == swRowB && ocB < swColB)))) { swB = oB; swRowB = orB; swColB = ocB; }
            }
            if (lane == 0) {
                const int sa = swA > a.dec_sub ? (swA - a.dec_sub) >> 2 : 0, sb = swB > a.dec_sub ? (swB - a.dec_sub) >> 2 : 0;
                a.scores[pa] = sa;
                if (a.end_rc) { a.end_rc[2 * pa] = sa > 0 ? swRowA : 0; a.end_rc[2 * pa + 1] = sa > 0 ? swColA : 0; }
                if (pb >= 0) {
                    a.scores[pb] = sb;
                    if (a.end_rc) { a.end_rc[2 * pb] = sb > 0 ? swRowB : 0; a.end_rc[2 * pb + 1] = sb > 0 ? swColB : 0; }
                }
            }
            continue;
        }
        // pairs with an empty sequence: the score is a pure border cell, nothing was captured above
        if (lane == 0) {
            if (QA == 0 || RA == 0) { const int n = QA + RA; a.scores[pa] = (((n == 0 ? a.b0 : a.b1 + a.bstep * n) - a.dec_sub) >> 2) + a.dec_add; }
            if (pb >= 0 && (QB == 0 || RB == 0)) { const int n = QB + RB; a.scores[pb] = (((n == 0 ? a.b0 : a.b1 + a.bstep * n) - a.dec_sub) >> 2) + a.dec_add; }
            if (a.end_rc) {
                a.end_rc[2 * pa] = QA; a.end_rc[2 * pa + 1] = RA;
                if (pb >= 0) { a.end_rc[2 * pb] = QB; a.end_rc[2 * pb + 1] = RB; }
            }
        }
    }
}

// ---- backtrack over the slab written above ---------------------------------------------------------------------
// Walk rules: LNW c++/LinearNeedlemanWunsch.cpp:137-223, ANW c++/AffineNeedlemanWunsch.cpp:242-403 (3-state walk, then
// pad rows as deletions, then columns as insertions :366-378).  One thread walks one pair.
struct PwBtArgs {
    const uint8_t* blob;
    const dpx_seq_pair* pairs;
    const int32_t* order;
    int first, count;
    int K;
    const int32_t* scores; const int32_t* end_rc;   // Smith-Waterman: start cells
    const uint32_t* tb;
    unsigned long long tb_stride;        // words per slot
    char* strings;
    const unsigned long long* str_off;
    int32_t* str_start;
};

template <int ALGO, int K, bool PACKED>
__global__ void __launch_bounds__(128) pw_bt_kernel(const PwBtArgs a) {
    constexpr bool AFF = (ALGO == DPX_ALGO_ANW);
    constexpr int CB = AFF ? 4 : 2;
    constexpr int CPH = 16 / CB;
    constexpr int W = K / CPH;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const int pos = a.first + t;
    const int pid = a.order ? a.order[pos] : pos;
    const dpx_seq_pair pr = a.pairs[pid];
    const int R = pr.referenceSize, Q = pr.querySize;
    int Rw = R;
    if (PACKED && (t ^ 1) < a.count) { const int po = a.order ? a.order[a.first + (t ^ 1)] : a.first + (t ^ 1); Rw = max(Rw, a.pairs[po].referenceSize); }
    const PwGeom geo = PwGeom::make(K, CB);
    const int nsteps2 = geo.nsteps(Rw);
    const int half = PACKED ? (t & 1) * 16 : 0;
    const uint8_t* __restrict__ ref = a.blob + pr.referenceIdx;
    const uint8_t* __restrict__ qry = a.blob + pr.queryIdx;
    const uint32_t* __restrict__ tb = a.tb + (unsigned long long)(PACKED ? (t >> 1) : t) * a.tb_stride;

    const size_t F = (size_t)Q + R + 1;
    char* __restrict__ o0 = a.strings + a.str_off[pid];
    char* __restrict__ o1 = o0 + F;
    char* __restrict__ o2 = o1 + F;
    long long p = (long long)F - 1;
    o0[p] = 0; o1[p] = 0; o2[p] = 0;

    auto code = [&](int i, int j) -> uint32_t {
        const int ii = i - 1;
        const int ps = ii / (32 * K), ln = (ii / K) & 31, r = ii % K;
        const int s = j - 1 + ln;
        const uint32_t w = __ldg(tb + (((size_t)ps * nsteps2 + s) * W + r / CPH) * 32 + ln);
        return (w >> (half + (CPH - 1 - r % CPH) * CB)) & ((1u << CB) - 1u);
    };
    auto emit_diag = [&](int i, int j) { const char rc = ref[j - 1], qc = qry[i - 1]; --p; o0[p] = rc; o1[p] = (rc == qc) ? '*' : '|'; o2[p] = qc; };
    auto emit_up   = [&](int i)        { --p; o0[p] = '_'; o1[p] = ' '; o2[p] = qry[i - 1]; };
    auto emit_left = [&](int j)        { --p; o0[p] = ref[j - 1]; o1[p] = ' '; o2[p] = '_'; };

    int i = Q, j = R;
    if (ALGO == DPX_ALGO_LSW) {
        // c++/LinearSmithWaterman.cpp:163-226: from the end cell apply the cell's direction, stop when the next cell's H is 0
        if (a.scores[pid] > 0) {
            i = a.end_rc[2 * pid]; j = a.end_rc[2 * pid + 1];
            uint32_t c = code(i, j);
            while (c != PW_SW_STOP) {
                if (c == PW_SW_DIAG) { emit_diag(i, j); --i; --j; }
                else if (c == PW_SW_UP) { emit_up(i); --i; }
                else { emit_left(j); --j; }
                if (i == 0 || j == 0) break;
                c = code(i, j);
            }
        }
    } else if (!AFF) {
        while (i != 0 || j != 0) {
            const uint32_t c = (i == 0) ? PW_LEFT : (j == 0) ? PW_UP : code(i, j);
            if (c == PW_DIAG) { emit_diag(i, j); --i; --j; }
            else if (c == PW_UP) { emit_up(i); --i; }
            else { emit_left(j); --j; }
        }
    } else {
        int state = 0;                   // 0 SCORING, 1 INSERTION, 2 DELETION
        while (i != 0 && j != 0) {
            const uint32_t c = code(i, j);
            if (state == 0) {
                const uint32_t d = c & 3u;
                if (d == PW_DIAG) { emit_diag(i, j); --i; --j; }
                else if (d == PW_UP) state = 2;
                else state = 1;
            } else if (state == 1) {
                state = (c & PW_IOPEN) ? 0 : 1;
                emit_left(j); --j;
            } else {
                state = (c & PW_DOPEN) ? 0 : 2;
                emit_up(i); --i;
            }
        }
        while (i > 0) { emit_up(i); --i; }
        while (j > 0) { emit_left(j); --j; }
    }
    a.str_start[pid] = (int32_t)p;
}

}  // namespace dpx
