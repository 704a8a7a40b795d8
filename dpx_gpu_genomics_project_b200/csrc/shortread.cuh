// shortread.cuh — LinearSmithWaterman score (+ end cell) for batches of short pairs: packed int16x2 DPX.
//
// Reference semantics: c++/LinearSmithWaterman.cpp:70-114 (recurrence, ReLU) and :145-157 (end cell =
// first strict maximum in row-major order).  No traceback here (the wavefront kernels do that).
//
// Mapping.  A group of G lanes aligns TWO pairs at once, pair A in the low and pair B in the high int16
// half of every register (both padded to common dimensions with never-matching pad symbols).  Lane g of
// the group keeps K consecutive query rows in registers and sweeps the reference columns one step behind
// lane g-1 (anti-diagonal wavefront of K-row blocks); the bottom row of a lane reaches the next lane with
// one __shfl_up per column step, i.e. one shuffle per K cell-pairs.  G*K rows are one pass; longer queries
// take several passes with the last lane's row handed over through a per-group shared-memory row.
//
// Arithmetic (per cell-pair; measured pipe model, DESIGN.md section 3: ALU pipe 2 clocks per warp instruction, plain adds 1, IMAD 2
// on the FMA pipe and about one clock next to packed DPX work):
//   s   = prmt(ta[r], tb[r], rsel)     per-row tables: byte c of ta / tb = (score - gap) of this row's base against reference code c
//                                      for pair A / pair B; the column's selector picks both and sign-extends them (ALU)
//   [WIDE: alphabets of 5..8 symbols -- the reference's data sets use '0'..'4' -- need all eight bytes of (ta, tb) for ONE pair:
//    a slot then holds one pair (its scores ride in the low halves, the high halves idle) and sequences are read as byte codes]
//   e   = __viaddmax_s16x2(diag, s, left) max(diag + s, left)   — independent of the row above (VIADDMNMX.S16x2)
//   h   = __vimax3_s16x2(e, up, B2)       ReLU against the bias B (= zero)                     (VIMNMX3.S16x2)
//   hg  = h + G2                       plain 32-bit add: all values carry a bias B >= -gap, so the low half
//                                      always carries into the high half and G2's high half is gap-1  (VIADD: 1 clock; as an IMAD
//                                      the step took 0.9 clocks per row longer, tools/sr_rowmix_bench.cu)
//   key = h * 2^k + code               one IMAD on the FMA pipe: the (positive, biased) score moves up k bits in
//                                      both halves (UNSIGNED 16-bit keys: (Hmax + B) << k < 65536); the low k bits are
//                                      [upper row of the row pair : 1][2^(k-1)-1 - (step mod 2^(k-1)) : k-1]; the two code
//                                      words count down with one IMAD each per step
//   best= vimax3.u16x2(best, key_r, key_r+1)   one VIMNMX3 per TWO rows: highest score, then upper row, then
//                                      earliest step
//   Every 2^(k-1) steps (warp-uniform; 64 steps for 150 bp reads at match 3) the row-pair registers are folded into one
//   32-bit key per pair and lane, (score | 31-row | 255-block | code): higher score, then smaller row, then earlier step —
//   exactly the reference's first-strict-max-in-row-major rule; lanes / passes are merged with the same order.  The fold
//   spreads the three fields with masks (ALU) and multiplies by run-time powers of two (FMA pipe), two candidates per VIMNMX3.U32.
// => 3.5 two-clock ALU-pipe instructions + 1 VIADD + 1 IMAD per cell-pair (2 cells) = 8.9 clocks measured; one IADD3 per column step.
// Registers hold hg = H + gap + B ("already gapped"), which is what the right and lower neighbours need;
// the diagonal neighbour wants H, so the table holds score - gap.
#pragma once
#include "common.cuh"

namespace dpx {

struct SrArgs {
    const uint32_t* packed;              // 2-bit packed sequences (pack.cuh)
    const uint8_t* codes;                // WIDE: byte codes 0..7, indexed like the blob (referenceIdx / queryIdx)
    const unsigned long long* pk_off;    // [n_pairs] word offsets, or null: pair p starts at p * pk_stride
    unsigned long long pk_stride;
    const dpx_seq_pair* pairs;
    const int32_t* order;                // schedule (nullable = identity); slot s = schedule positions 2s, 2s+1
    int n_pairs, n_slots;
    uint32_t ms_byte, xs_byte;           // (match - gap) & 0xff, (mismatch - gap) & 0xff
    uint32_t B2, Bg2, G2;                // packed bias, bias + gap, and the add constant ((gap-1)<<16 | gap&0xffff)
    int B;
    uint32_t one;                        // 1 (run-time constant for the FMA-pipe adds)
    uint32_t kmul;                       // 1 << kbits, passed as data so the shift stays an FMA-pipe IMAD
    int kbits;                           // position bits per int16 key: (Hmax + B) << kbits must stay < 32768
    int32_t* scores;
    int32_t* end_rc;                     // nullable
    unsigned int* counter;
    int bnd_stride, rsel_stride;         // per-group shared-memory strides (uint32 / uint16 entries)
};

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));   // generic mode: nibble msb = replicate sign
    return d;
}

// Integer ops pinned to the FMA pipe: mad.lo with a run-time multiplier of 1 (`one` comes from the kernel
// arguments, so ptxas cannot fold it into an ALU-pipe add / select).  The ALU pipe (DPX, PRMT) is the
// bottleneck of this kernel; the FMA pipe issues in the other half of the SMSP's cycles.
__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t one, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

// A plain 32-bit add that stays one (tools/sr_rowmix_bench.cu: next to the half-rate DPX / PRMT instructions a VIADD costs about one
// clock per row less than the same add as an IMAD -- the two half-rate pipes do not overlap perfectly, a full-rate add slips in between).
__device__ __forceinline__ uint32_t alu_add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t fma_mad(uint32_t a, uint32_t m, uint32_t b) {   // a * m + b, m a run-time value: stays an IMAD
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) {       // PTX shl clamps the amount: n >= 32 gives 0
    uint32_t d;
    asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(n));
    return d;
}

__device__ __forceinline__ uint32_t get2(const uint32_t* __restrict__ w, int k) { return (w[k >> 4] >> (2 * (k & 15))) & 3u; }

template <int G, int K, bool TRACK, bool WIDE>
__global__ void __launch_bounds__(128, 4) sr_lsw_kernel(const SrArgs a) {
    extern __shared__ uint32_t sr_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G;                    // groups per warp
    constexpr int GPB = 128 / G;                   // groups per block
    const int lane = threadIdx.x & 31, gl = lane % G, gw = lane / G;
    const int gib = threadIdx.x / G;
    uint32_t* __restrict__ bnd = sr_smem + (size_t)gib * a.bnd_stride + G;        // bnd[j], j >= -G + 2 (the last lane's early steps land in front)
    uint16_t* __restrict__ rsel = reinterpret_cast<uint16_t*>(sr_smem + (size_t)GPB * a.bnd_stride) + (size_t)gib * a.rsel_stride;
    const uint32_t B2 = a.B2, Bg2 = a.Bg2, G2 = a.G2, one = a.one;
    const uint32_t xs4 = a.xs_byte * 0x01010101u, dms = (a.ms_byte ^ a.xs_byte) & 0xffu;      // xor: no byte carries, any signs
    const int kmask = (1 << a.kbits) - 1;             // all position bits of an int16 key
    const int smask = (1 << (a.kbits - 1)) - 1;       // step-code bits (the bit above them marks the upper row of a pair)
    const uint32_t kmul = a.kmul;

    for (;;) {
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(a.counter, (unsigned)GPW);
        base = __shfl_sync(FULL, base, 0);
        if (base >= a.n_slots) break;
        const int slot = base + gw;
        int pa = -1, pb = -1, RA = 0, QA = 0, RB = 0, QB = 0;
        const uint32_t *refA = a.packed, *qryA = a.packed, *refB = a.packed, *qryB = a.packed;
        const uint8_t *crefA = a.codes, *cqryA = a.codes;            // WIDE
        if (WIDE && slot < a.n_slots) {
            pa = a.order ? a.order[slot] : slot;                      // one pair per slot
            const dpx_seq_pair p = a.pairs[pa];
            RA = p.referenceSize; QA = p.querySize;
            crefA = a.codes + p.referenceIdx; cqryA = a.codes + p.queryIdx;
        }
        if (!WIDE && slot < a.n_slots) {
            pa = a.order ? a.order[2 * slot] : 2 * slot;
            const dpx_seq_pair p = a.pairs[pa];
            RA = p.referenceSize; QA = p.querySize;
            refA = a.packed + (a.pk_off ? a.pk_off[pa] : (unsigned long long)pa * a.pk_stride); qryA = refA + ((RA + 15) >> 4);
            if (2 * slot + 1 < a.n_pairs) {
                pb = a.order ? a.order[2 * slot + 1] : 2 * slot + 1;
                const dpx_seq_pair q = a.pairs[pb];
                RB = q.referenceSize; QB = q.querySize;
                refB = a.packed + (a.pk_off ? a.pk_off[pb] : (unsigned long long)pb * a.pk_stride); qryB = refB + ((RB + 15) >> 4);
            }
        }
        const int Rw = __reduce_max_sync(FULL, max(RA, RB));
        const int Qw = __reduce_max_sync(FULL, max(QA, QB));
        const int passes = (Qw + G * K - 1) / (G * K);
        const int nsteps2 = (Rw + G - 1 + 1) & ~1;          // column steps incl. pipeline drain, rounded to even

        // ---- per-group column table: entry e <-> column j = e - (G-1); pads never match ---------------
        __syncwarp();
        // selector of the per-row score tables (below): nibble 0 = reference base of pair A (bytes 0..3 of the row's table pair),
        // nibble 2 = 4 + base of pair B (bytes 4..7); nibbles 1 / 3 = the same index | 8 = "replicate that byte's sign", i.e. the
        // int16 sign extension.  Pad columns select the sign of byte 0 / 4: 0 or -1, below any real score.
        // Built one packed word (16 bases) per lane at a time: base k of the reference is entry e = k + G.
        for (int e = gl; e < G; e += G) rsel[e] = (uint16_t)0xcc88u;                       // columns j <= 0
        if (WIDE) {
            // one pair, codes 0..7: nibble 0 = code (byte of the row's 8-byte table ta:tb), nibble 1 = code | 8 (its sign); the high
            // half repeats a pad (sign of byte 4 -- an entry >= 0 would do as well: the high halves carry nothing)
            for (int k = gl; k < Rw + G + 2; k += G) {               // lane 0 runs G - 1 pad columns past the last real one
                const uint32_t nA = (k < RA) ? (uint32_t)(crefA[k] & 7u) : 8u;
                rsel[k + G] = (uint16_t)((k < RA) ? (nA | ((nA | 8u) << 4) | 0xcc00u) : 0xcc88u);
            }
        }
        if (!WIDE) {
            // words whose 16 bases exist in both pairs: every lane builds the two entries (2 gl, 2 gl + 1) of each word.  A byte
            // of an entry is 0x11 * base + 0x80 (pair A) / + 0xc4 (pair B): the two 4-bit fields of a word are spread over four
            // bytes by one multiplication (x + (x << 14), masked), the constant 0x11 by another.
            const int nfull = min(RA, RB) >> 4;
            uint32_t* __restrict__ rsel32 = reinterpret_cast<uint32_t*>(rsel + G) + gl;
            #pragma unroll 4
            for (int c = 0; c < nfull; ++c) {
                const uint32_t x = (refA[c] >> (4 * gl)) & 0xfu, y = (refB[c] >> (4 * gl)) & 0xfu;
                const uint32_t z = fma_mad(y, a.one << 8, x);
                const uint32_t z2 = fma_mad(z, a.one * 0x4001u, 0u) & 0x03030303u;
                rsel32[8 * c] = fma_mad(z2, a.one * 0x11u, 0xc480c480u);
            }
            // the rest, pads included, one entry at a time
            for (int k = 16 * nfull + gl; k < Rw + G + 2; k += G) {
                const uint32_t nA = (k < RA) ? get2(refA, k) : 8u;
                const uint32_t nB = (k < RB) ? (4u + get2(refB, k)) : 12u;
                rsel[k + G] = (uint16_t)(nA | ((nA | 8u) << 4) | (nB << 8) | ((nB | 8u) << 12));
            }
        }
        for (int e = gl; e < Rw + G + 2; e += G) bnd[e] = Bg2;                      // pass 0 reads the bias everywhere
        __syncwarp();

        int bestA = 0, rowA = 0, colA = 0, bestB = 0, rowB = 0, colB = 0;

        for (int p = 0; p < passes; ++p) {
            const int i0 = p * G * K + gl * K;                  // matrix row of register row r is i0 + r + 1
            uint32_t ta[K], tb[K], hgA[K], hgB[K], best[(K + 1) / 2];   // ta / tb: (score - gap) of this row's base against bases 0..3
            uint32_t laneKeyA = 0, laneKeyB = 0;                 // score<<(k+12) | (31-row)<<(k+7) | (255-blk)<<(k-1) | code
            #pragma unroll
            for (int r = 0; r < K; ++r) {
                const int i = i0 + r;                            // 0-based query index
                if (WIDE) {
                    const uint32_t qa = (i < QA) ? (uint32_t)(cqryA[i] & 7u) : 8u;           // 8: a pad row matches nothing
                    ta[r] = xs4 ^ shl_clamp(dms, (qa < 4u) ? 8u * qa : 32u);
                    tb[r] = xs4 ^ shl_clamp(dms, (qa >= 4u && qa < 8u) ? 8u * (qa - 4u) : 32u);
                } else {
                    const uint32_t qa = (i < QA) ? get2(qryA, i) : 4u;
                    const uint32_t qb = (i < QB) ? get2(qryB, i) : 4u;
                    ta[r] = xs4 ^ shl_clamp(dms, (qa < 4u) ? 8u * qa : 32u);
                    tb[r] = xs4 ^ shl_clamp(dms, (qb < 4u) ? 8u * qb : 32u);
                }
                hgA[r] = Bg2; hgB[r] = Bg2;
            }
            #pragma unroll
            for (int r = 0; r < (K + 1) / 2; ++r) best[r] = 0;
            uint32_t best2 = B2;                                 // !TRACK: one running max for all rows
            uint32_t bot = Bg2, topprev = Bg2;
            const bool store_bot = (gl == G - 1) && (p + 1 < passes);       // j < 1 lands in the slack in front of the row
            const uint32_t csdec = 0xfffefffFu;                             // -1 in both halves

#define DPX_SR_STEP(S, OLD, NEW)                                                                         \
            {                                                                                            \
                const int j = (S) - gl + 1;                                                              \
                uint32_t top = __shfl_up_sync(FULL, bot, 1, G);                                          \
                if (gl == 0) top = bnd[j];                                                               \
                const uint32_t rs = rsel[(S) - gl + G];                                                  \
                uint32_t keyprev = 0;                                                                    \
                uint32_t upg = top, diag = topprev;                                                      \
                topprev = top;                                                                           \
                _Pragma("unroll")                                                                        \
                for (int r = 0; r < K; ++r) {                                                            \
                    const uint32_t sc = prmt_b32(ta[r], tb[r], rs);          /* both pairs' scores, sign-extended */ \
                    const uint32_t e = __viaddmax_s16x2(diag, sc, OLD[r]);   /* off the row chain */     \
                    const uint32_t h = __vimax3_s16x2(e, upg, B2);           /* chain: h -> hg -> h */   \
                    diag = OLD[r];                                                                       \
                    NEW[r] = alu_add(h, G2);                                                              \
                    upg = NEW[r];                                                                        \
                    if (TRACK) {                                                                         \
                        const uint32_t key = fma_add(h, kmul, (r & 1) ? csL : csU);  /* IMAD, FMA pipe */\
                        if (r & 1) best[r >> 1] = __vimax3_u16x2(best[r >> 1], keyprev, key);            \
                        else if (r == K - 1) best[r >> 1] = __vmaxu2(best[r >> 1], key);                 \
                        keyprev = key;                                                                   \
                    } else {                                                                             \
                        if (r & 1) best2 = __vimax3_s16x2(best2, keyprev, h);                            \
                        else if (r == K - 1) best2 = __vmaxs2(best2, h);                                 \
                        keyprev = h;                                                                     \
                    }                                                                                    \
                }                                                                                        \
                bot = NEW[K - 1];                                                                        \
                if (TRACK) { csL = fma_add(csL, one, csdec); csU = fma_add(csU, one, csdec); }           \
                if (store_bot) bnd[j] = bot;                                                             \
            }

            const int bs = smask + 1;                            // steps per position block (even)
            // field multipliers of the fold, run-time values so that the spreading stays on the FMA pipe
            const uint32_t m12 = one << 12, m8 = one << 8, nrow = 0u - (one << (a.kbits + 8));
            const uint32_t smA = 0xffffu & ~(uint32_t)kmask, rbit = (uint32_t)bs;
            for (int blk = 0; blk * bs < nsteps2; ++blk) {
                const int s_end = min((blk + 1) * bs, nsteps2);
                uint32_t csL = (uint32_t)smask * 0x00010001u;    // step code of the lower row, counts down; the upper row's carries the row bit
                uint32_t csU = csL + (uint32_t)bs * 0x00010001u;
                #pragma unroll 1
                for (int s = blk * bs; s < s_end; s += 2) {
                    DPX_SR_STEP(s, hgA, hgB)
                    DPX_SR_STEP(s + 1, hgB, hgA)
                }
                if (TRACK) {
                    // fold the row-pair registers into the lane keys (warp-uniform point).  lane key =
                    // score << (k+12) | (31 - row) << (k+7) | (255 - blk) << (k-1) | step code
                    const uint32_t tag0 = ((uint32_t)(255 - blk) << (a.kbits - 1)) | (30u << (a.kbits + 7));
                    uint32_t cA[(K + 1) / 2], cB[(K + 1) / 2];
                    #pragma unroll
                    for (int q = 0; q < (K + 1) / 2; ++q) {
                        const uint32_t tag = fma_mad((uint32_t)q, nrow, tag0);               // 30 - 2q: lower row of the pair; the row bit adds 1
                        const uint32_t x = best[q], y = x >> 16;
                        cA[q] = fma_mad(x & smA, m12, fma_mad(x & rbit, m8, fma_add(x & (uint32_t)smask, one, tag)));
                        cB[q] = fma_mad(y & smA, m12, fma_mad(y & rbit, m8, fma_add(y & (uint32_t)smask, one, tag)));
                    }
                    #pragma unroll
                    for (int q = 0; q + 1 < (K + 1) / 2; q += 2) {
                        laneKeyA = __vimax3_u32(laneKeyA, cA[q], cA[q + 1]);
                        laneKeyB = __vimax3_u32(laneKeyB, cB[q], cB[q + 1]);
                    }
                    if (((K + 1) / 2) & 1) {
                        laneKeyA = max(laneKeyA, cA[(K + 1) / 2 - 1]);
                        laneKeyB = max(laneKeyB, cB[(K + 1) / 2 - 1]);
                    }
                    // (best[] runs on: a key that survives from an earlier block was folded there with a better block tag)
                }
            }
#undef DPX_SR_STEP

            if (TRACK) {
                // decode: step of the maximum = block * bs + (bs-1 - code); column = step - lane + 1
                const int hA = (int)(laneKeyA >> (12 + a.kbits)) - a.B;
                const int hB = (int)(laneKeyB >> (12 + a.kbits)) - a.B;
                if (hA > bestA) {
                    bestA = hA; rowA = i0 + (31 - (int)((laneKeyA >> (7 + a.kbits)) & 31u)) + 1;
                    colA = (255 - (int)((laneKeyA >> (a.kbits - 1)) & 255u)) * bs + (smask - (int)(laneKeyA & (uint32_t)smask)) - gl + 1;
                }
                if (hB > bestB) {
                    bestB = hB; rowB = i0 + (31 - (int)((laneKeyB >> (7 + a.kbits)) & 31u)) + 1;
                    colB = (255 - (int)((laneKeyB >> (a.kbits - 1)) & 255u)) * bs + (smask - (int)(laneKeyB & (uint32_t)smask)) - gl + 1;
                }
            } else {
                bestA = max(bestA, (int)(short)(best2 & 0xffffu) - a.B);
                bestB = max(bestB, (int)(short)(best2 >> 16) - a.B);
            }
            __syncwarp();
        }

        // ---- reduce over the lanes of the group: higher score, then smaller row ---------------------------
        #pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            const int sA = __shfl_xor_sync(FULL, bestA, off), rA_ = __shfl_xor_sync(FULL, rowA, off), cA_ = __shfl_xor_sync(FULL, colA, off);
            const int sB = __shfl_xor_sync(FULL, bestB, off), rB_ = __shfl_xor_sync(FULL, rowB, off), cB_ = __shfl_xor_sync(FULL, colB, off);
            if (sA > bestA || (sA == bestA && rA_ < rowA)) { bestA = sA; rowA = rA_; colA = cA_; }
            if (sB > bestB || (sB == bestB && rB_ < rowB)) { bestB = sB; rowB = rB_; colB = cB_; }
        }
        if (gl == 0) {
            if (pa >= 0) { a.scores[pa] = bestA; if (TRACK && a.end_rc) { a.end_rc[2 * pa] = rowA; a.end_rc[2 * pa + 1] = colA; } }
            if (pb >= 0) { a.scores[pb] = bestB; if (TRACK && a.end_rc) { a.end_rc[2 * pb] = rowB; a.end_rc[2 * pb + 1] = colB; } }
        }
    }
}

}  // namespace dpx
