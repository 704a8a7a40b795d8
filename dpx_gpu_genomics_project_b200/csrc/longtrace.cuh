// longtrace.cuh — alignment strings of ONE very long LinearSmithWaterman pair from CHECKPOINTS instead of a traceback matrix.
//
// A 1 Mbp x 1 Mbp direction matrix is 250 GB even at 2 bits per cell (the reference's 8 B per cell: 8 TB,
// c++/LinearSmithWaterman.cpp:32-45), so the forward kernel (longpair.cuh) keeps none.  In checkpoint mode it keeps a GRID of
// H instead: every warp stores the column it reads from its left neighbour (32 rows per coalesced store), i.e. H[i][c * TW]
// for every row i and every block boundary c * TW, and every lane stores its columns of every TH-th row, H[r * TH][j]
// (4 B * R * Q * (1 / TW + 1 / TH): 12 GB at 1 Mbp x 1 Mbp with 512 x 1024 tiles).  Together they are the top row and the
// left column of every TH x TW tile of the matrix.
//
// The walk (c++/LinearSmithWaterman.cpp:160-226 restated) then needs the directions of one tile at a time: starting from the
// end cell, re-fill the tile that holds the current cell from its two borders with the reference's direction rule
// (UP if up == H, else LEFT if left == H, else DIAG; STOP iff H == 0, :104-108 and :222), follow the directions to the tile's
// edge, move to the neighbouring tile.  At most Q / TH + R / TW + 1 tiles are ever filled: 2 * 10^9 cells for the 10^12-cell
// matrix.  Every H the walk looks at is the forward pass's own value, so the strings are those the reference's full-matrix
// walk would print (tests/test_gpu_longtrace.py: bit-exact against the oracle's full-matrix backtrack up to 12 kbp at every tile size).
//
// Tiles do not depend on the walk, only on their borders, so they are filled AHEAD of it, many at a time: the host predicts
// the tiles the path will cross (a band around the line through the current cell along the walk's recent direction, diagonal
// at first) and three kernels run per round:
//   fill   one block per predicted tile.  Thread t owns four columns of the tile and sweeps their rows with a skew of t
//          (anti-diagonal wavefront, one __syncthreads per step); directions are packed 16 rows per word into the tile's slot
//          (global memory, L2 resident).  Beside H every cell carries WHERE A WALK STARTING THERE LEAVES THE TILE and after how
//          many moves — taken over from the neighbour its own direction points at, so it costs a few selects per cell — and
//          the tile's last row and last column (the only cells a walk can enter through) keep that as the tile's transfer table.
//   chain  one warp hops from tile to tile through the transfer tables: entry cell -> (exit cell, moves).  No direction is
//          read; a hop is two dependent loads.  It notes, per tile, the entry cell and the offset of its moves in the output,
//          and stops when the walk ends or enters a tile that was not predicted (then the next round predicts from there).
//   emit   one block per tile the chain went through, all in parallel: stage the slot into shared memory, thread 0 follows
//          the directions from the entry cell to the tile's edge and writes the three lines at the offset the chain gave it.
// A wrong prediction costs a round, never a wrong result.
#pragma once
#include "common.cuh"

namespace dpx {

struct LongBtArgs {
    const uint8_t* ref;                  // raw reference bytes (compared for equality, as the reference does)
    const uint8_t* qry;
    long long ie, je;                    // end cell (1-based matrix row / column): the walk never looks right of or below it
    int match, mismatch, gap;
    int TH, TW;                          // tile height = row-checkpoint spacing, tile width = column-checkpoint spacing
    const int32_t* colck;                // colck[c * col_stride + i] = H[i][(c + 1) * TW], rows 1..Q (checkpoint dumps of longpair.cuh)
    long long col_stride;
    const int32_t* rowck;                // rowck[r * row_stride + j] = H[(r + 1) * TH][j], columns 1..R (the host passes base - 1)
    long long row_stride;
    const unsigned long long* tiles;     // this round's tiles: (tile row << 32) | tile column; slot b holds tile b
    int ntiles;
    uint32_t* slots;                     // [ntiles][ceil(TH / 16)][TWp] direction words
    uint32_t* edges;                     // [ntiles] transfer tables: exit[TWp] moves[TWp] of the last row, exit[THp] moves[THp] of the last column
    int4* segs;                          // [ntiles] chain -> emit: {slot, entry row, entry column, moves} of every tile walked this round
    long long* seg_off;                  // [ntiles] characters written before that tile
    uint8_t* out;                        // three lines of `cap` bytes each (REF, REL, QRY), written from the back
    long long cap;
    long long* state;                    // [0] row i, [1] column j, [2] characters written, [3] done, [4] tiles walked, [5] segments of this round, [6] error
};

constexpr int LONG_BT_CPT = 4;           // tile columns per thread of the fill kernel
constexpr uint32_t LONG_BT_STOP = 0x80000000u;
// exit cell of a walk, relative to the tile: (row + 1) << 11 | (column + 1); row + 1 == 0 / column + 1 == 0: the cell above / left of the tile
DPX_HD uint32_t long_bt_pack(int row_p1, int col_p1) { return ((uint32_t)row_p1 << 11) | (uint32_t)col_p1; }

// a slot holds a tile's directions, 16 rows of one column per word: [ceil(TH / 16)][TWp]
DPX_HD size_t long_bt_slot_words(int TH, int TW) { return (size_t)((TH + 15) / 16) * (size_t)((TW + 31) & ~31); }
DPX_HD size_t long_bt_edge_words(int TH, int TW) { return 2 * (size_t)((TW + 31) & ~31) + 2 * (size_t)((TH + 31) & ~31); }
// emit kernel's shared memory: the slot | uint8 sq[TH] | uint8 sr[TWp]
DPX_HD size_t long_bt_walk_smem(int TH, int TW) { return long_bt_slot_words(TH, TW) * 4 + (size_t)((TH + 15) & ~15) + (size_t)((TW + 31) & ~31); }
// fill kernel's shared memory: int hbuf[3][2][NT] (H, exit, moves) | int left[TH + 1] | uint8 sq[TH]
DPX_HD size_t long_bt_fill_smem(int TH, int NT) { return (size_t)6 * NT * 4 + (size_t)(TH + 1) * 4 + (size_t)((TH + 15) & ~15); }

__global__ void __launch_bounds__(256) long_tile_fill_kernel(const LongBtArgs a) {
    constexpr int CPT = LONG_BT_CPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int TH = a.TH, TW = a.TW, TWp = (TW + 31) & ~31, THp = (TH + 31) & ~31;
    int* hbuf = reinterpret_cast<int*>(smem_raw);                       // [2][NT] H of the thread's last column
    uint32_t* ebuf = reinterpret_cast<uint32_t*>(hbuf + 2 * NT);        // [2][NT] its exit cell
    int* lbuf = reinterpret_cast<int*>(ebuf + 2 * NT);                  // [2][NT] its moves
    int* sleft = lbuf + 2 * NT;
    uint8_t* sq = reinterpret_cast<uint8_t*>(sleft + TH + 1);
    const int g = a.gap;
    const unsigned long long key = a.tiles[blockIdx.x];
    const long long tr = (long long)(key >> 32), tc = (long long)(key & 0xffffffffull);
    const long long r0 = tr * TH, c0 = tc * TW;
    const int h = (int)min((long long)TH, a.ie - r0), w = (int)min((long long)TW, a.je - c0);     // tiles on the end cell's row / column are cut there
    const int wt = (w + CPT - 1) / CPT;                                 // threads with at least one column
    const int klast = (w - 1) & (CPT - 1);                              // the tile's last column inside thread wt - 1
    uint32_t* __restrict__ slot = a.slots + (size_t)blockIdx.x * long_bt_slot_words(TH, TW);
    uint32_t* __restrict__ edge = a.edges + (size_t)blockIdx.x * long_bt_edge_words(TH, TW);
    // ---- borders and sequence slices of the tile (row 0 and column 0 of the matrix are 0 and have no checkpoint) ------------
    for (int k = tid; k < h; k += NT) sq[k] = a.qry[r0 + k];
    for (int k = tid; k <= h; k += NT)
        sleft[k] = (tc > 0 && r0 + k > 0) ? a.colck[(tc - 1) * a.col_stride + r0 + k] : 0;            // H[r0 + k][c0]
    uint32_t rcv[CPT], upE[CPT];
    int up[CPT], upL[CPT];
    int diag0 = 0;                                                      // H[r - 1][column left of this thread's first]
    #pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = tid * CPT + k;
        rcv[k] = c < w ? (uint32_t)a.ref[c0 + c] : 0x100u;              // columns right of the cut never match; nobody reads them
        up[k] = (c < w && tr > 0) ? a.rowck[(tr - 1) * a.row_stride + c0 + 1 + c] : 0;              // H[r0][c0 + 1 + c]
        upE[k] = long_bt_pack(0, c + 1); upL[k] = 0;                    // a walk that goes up from row 0 leaves through the cell above: 1 move
    }
    if (tid < wt && tr > 0 && c0 + (long long)tid * CPT > 0) diag0 = a.rowck[(tr - 1) * a.row_stride + c0 + (long long)tid * CPT];
    uint32_t diag0E = long_bt_pack(0, tid * CPT);                       // ... or through the cell above-left
    int diag0L = 0;
    __syncthreads();
    // ---- skewed sweep: at step s thread t fills tile row s - t of its CPT columns --------------------------------------------
    uint32_t acc[CPT];
    #pragma unroll
    for (int k = 0; k < CPT; ++k) acc[k] = 0;
    const int nsteps = h + wt - 1;
    for (int s = 0; s < nsteps; ++s) {
        const int r = s - tid;
        if (tid < wt && r >= 0 && r < h) {
            const int pb = ((s - 1) & 1) * NT + tid - 1;
            const int left_in = tid == 0 ? sleft[r + 1] : hbuf[pb];
            const uint32_t left_inE = tid == 0 ? long_bt_pack(r + 1, 0) : ebuf[pb];               // column 0: left and diagonal leave through column -1
            const int left_inL = tid == 0 ? 0 : lbuf[pb];
            const uint32_t qc = sq[r];
            int left = left_in, diag = diag0, leftL = left_inL, diagL = diag0L;
            uint32_t leftE = left_inE, diagE = diag0E;
            const int sh = 2 * (r & 15);
            const uint32_t selfE = LONG_BT_STOP | long_bt_pack(r + 1, tid * CPT + 1);
            #pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int u = up[k];
                const int ug = u + g, lg = left + g;
                const int dg = diag + (qc == rcv[k] ? a.match : a.mismatch);
                const int v = __vimax3_s32_relu(ug, lg, dg);
                const bool is_up = ug == v, is_left = lg == v;
                const uint32_t code = v == 0 ? C_STOP : (is_up ? C_UP : (is_left ? C_LEFT : C_DIAG));
                acc[k] |= code << sh;
                // the walk from this cell: stops here (0 moves), or one move to the neighbour the direction names and on from there
                const uint32_t e = v == 0 ? selfE + k : (is_up ? upE[k] : (is_left ? leftE : diagE));
                const int l = v == 0 ? 0 : (is_up ? upL[k] : (is_left ? leftL : diagL)) + 1;
                diag = u; diagE = upE[k]; diagL = upL[k];
                up[k] = v; upE[k] = e; upL[k] = l;
                left = v; leftE = e; leftL = l;
            }
            diag0 = left_in; diag0E = left_inE; diag0L = left_inL;
            const int cb = (s & 1) * NT + tid;
            hbuf[cb] = left; ebuf[cb] = leftE; lbuf[cb] = leftL;
            if ((r & 15) == 15 || r == h - 1) {
                static_assert(CPT == 4, "one 16-byte store per thread");
                *reinterpret_cast<uint4*>(slot + (size_t)(r >> 4) * TWp + tid * CPT) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                #pragma unroll
                for (int k = 0; k < CPT; ++k) acc[k] = 0;
            }
            // transfer table: the last row (every thread, once) and the last column (thread wt - 1, every row)
            if (r == h - 1) {
                #pragma unroll
                for (int k = 0; k < CPT; ++k) { edge[tid * CPT + k] = upE[k]; edge[TWp + tid * CPT + k] = (uint32_t)upL[k]; }
            }
            if (tid == wt - 1) {
                const uint32_t e = klast == 0 ? upE[0] : klast == 1 ? upE[1] : klast == 2 ? upE[2] : upE[3];
                const int l = klast == 0 ? upL[0] : klast == 1 ? upL[1] : klast == 2 ? upL[2] : upL[3];
                edge[2 * TWp + r] = e; edge[2 * TWp + THp + r] = (uint32_t)l;
            }
        }
        __syncthreads();
    }
}

// One warp, every lane on the same scalar path (loads are broadcasts); the lanes only split up to search the tile list.
__global__ void __launch_bounds__(32) long_chain_kernel(const LongBtArgs a) {
    const int lane = threadIdx.x;
    const int TH = a.TH, TW = a.TW, TWp = (TW + 31) & ~31, THp = (TH + 31) & ~31;
    long long i = a.state[0], j = a.state[1], n = a.state[2], done = a.state[3], tiles = a.state[4], err = a.state[6];
    int nseg = 0;
    while (!done && nseg < a.ntiles) {
        const long long tr = (i - 1) / TH, tc = (j - 1) / TW;
        const long long r0 = tr * TH, c0 = tc * TW;
        const unsigned long long key = ((unsigned long long)tr << 32) | (unsigned long long)tc;
        int b = -1;
        for (int base = 0; base < a.ntiles && b < 0; base += 32) {
            const bool hit = base + lane < a.ntiles && a.tiles[base + lane] == key;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) b = base + __ffs(m) - 1;
        }
        if (b < 0) break;                                              // not predicted: the round is over
        const int hf = (int)min((long long)TH, a.ie - r0), wf = (int)min((long long)TW, a.je - c0);
        const int x = (int)(i - r0) - 1, y = (int)(j - c0) - 1;       // entry cell inside the tile
        const uint32_t* __restrict__ edge = a.edges + (size_t)b * long_bt_edge_words(TH, TW);
        uint32_t e; int l;
        if (x == hf - 1) { e = __ldcg(edge + y); l = (int)__ldcg(edge + TWp + y); }
        else if (y == wf - 1) { e = __ldcg(edge + 2 * TWp + x); l = (int)__ldcg(edge + 2 * TWp + THp + x); }
        else { err = 1; break; }                                       // a walk can only enter through the last row or column
        if (lane == 0) { a.segs[nseg] = make_int4(b, x, y, l); a.seg_off[nseg] = n; }
        ++nseg; ++tiles;
        n += l;
        i = r0 + (long long)((e & ~LONG_BT_STOP) >> 11); j = c0 + (long long)(e & 0x7ffu);
        done = (e & LONG_BT_STOP) != 0 || i == 0 || j == 0;
    }
    if (lane == 0) { a.state[0] = i; a.state[1] = j; a.state[2] = n; a.state[3] = done; a.state[4] = tiles; a.state[5] = nseg; a.state[6] = err; }
}

__global__ void __launch_bounds__(256) long_emit_kernel(const LongBtArgs a) {
    if ((long long)blockIdx.x >= a.state[5]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int TH = a.TH, TW = a.TW, TWp = (TW + 31) & ~31;
    uint32_t* sdirs = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* sq = reinterpret_cast<uint8_t*>(sdirs + long_bt_slot_words(TH, TW));
    uint8_t* sr = sq + ((TH + 15) & ~15);
    const int4 seg = a.segs[blockIdx.x];
    const long long off = a.seg_off[blockIdx.x];
    const unsigned long long key = a.tiles[seg.x];
    const long long r0 = (long long)(key >> 32) * TH, c0 = (long long)(key & 0xffffffffull) * TW;
    const int h = seg.y + 1, w = seg.z + 1;                            // the part of the tile at or above-left of the entry cell
    // ---- stage the needed part of the slot and the two sequence slices ----------------------------------------------------------
    const uint32_t* __restrict__ slot = a.slots + (size_t)seg.x * long_bt_slot_words(TH, TW);
    const int wr = (h + 15) >> 4, wq = (w + 3) >> 2;                    // word rows, 16-byte groups per word row
    for (int x = tid; x < wr * wq; x += NT) {
        const int rr = x / wq, cq = x - rr * wq;
        reinterpret_cast<uint4*>(sdirs + (size_t)rr * TWp)[cq] = __ldcg(reinterpret_cast<const uint4*>(slot + (size_t)rr * TWp) + cq);
    }
    for (int k = tid; k < h; k += NT) sq[k] = a.qry[r0 + k];
    for (int k = tid; k < w; k += NT) sr[k] = a.ref[c0 + k];
    __syncthreads();
    // ---- walk inside the tile (c++/LinearSmithWaterman.cpp:160-226) ----------------------------------------------------------------
    if (tid == 0) {
        int ra = h, cb = w;
        uint8_t* p0 = a.out + (a.cap - 1) - off, *p1 = p0 + a.cap, *p2 = p1 + a.cap;
        uint8_t* const p0_start = p0;
        while (ra > 0 && cb > 0) {
            const int x = ra - 1, y = cb - 1;
            const uint32_t d = (sdirs[(x >> 4) * TWp + y] >> (2 * (x & 15))) & 3u;
            if (d == C_STOP) break;
            const uint8_t qi = sq[x], rj = sr[y];
            const bool dg = d == C_DIAG, upm = d == C_UP;
            *p0-- = upm ? (uint8_t)'_' : rj;                                       // REF line: '_' where the query base has no partner
            *p1-- = dg ? (qi == rj ? (uint8_t)'*' : (uint8_t)'|') : (uint8_t)' ';
            *p2-- = (dg || upm) ? qi : (uint8_t)'_';
            ra -= (dg || upm) ? 1 : 0;
            cb -= (dg || !upm) ? 1 : 0;
        }
        if ((long long)(p0_start - p0) != (long long)seg.w) a.state[6] = 2;       // the transfer table and the directions must agree
    }
}

}  // namespace dpx
