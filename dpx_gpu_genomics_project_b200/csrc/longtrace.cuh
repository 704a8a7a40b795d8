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
// walk would print (tests: bit-exact against the oracle's full-matrix backtrack up to 30 kbp, tile-size invariance).
//
// Tiles do not depend on the walk, only on their borders, so they are filled AHEAD of it, many at a time: the host predicts
// the tiles the path will cross (a band around the line through the current cell along the walk's recent direction, diagonal
// at first), one kernel fills them all, one block each (thread t owns four columns of the tile and sweeps their rows with a
// skew of t: anti-diagonal wavefront, one __syncthreads per step; directions packed 16 rows per word into the tile's slot in
// global memory, L2 resident), then a single block walks: it stages the current tile's slot into shared memory, thread 0
// follows the directions to the tile's edge writing the three lines from the back of their buffers, and goes on until the walk
// stops or steps onto a tile that was not predicted; then the next round predicts from there.  A wrong prediction costs a
// round, never a wrong result.
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
    uint8_t* out;                        // three lines of `cap` bytes each (REF, REL, QRY), written from the back
    long long cap;
    long long* state;                    // [0] current row i, [1] current column j, [2] characters written, [3] done, [4] tiles walked
};

constexpr int LONG_BT_CPT = 4;           // tile columns per thread of the fill kernel

// a slot holds a tile's directions, 16 rows of one column per word: [ceil(TH / 16)][TWp]
DPX_HD size_t long_bt_slot_words(int TH, int TW) { return (size_t)((TH + 15) / 16) * (size_t)((TW + 31) & ~31); }
// walker's shared memory: the slot | uint8 sq[TH] | uint8 sr[TWp]
DPX_HD size_t long_bt_walk_smem(int TH, int TW) { return long_bt_slot_words(TH, TW) * 4 + (size_t)((TH + 15) & ~15) + (size_t)((TW + 31) & ~31); }
// fill kernel's shared memory: int hbuf[2][NT] | int left[TH + 1] | uint8 sq[TH]
DPX_HD size_t long_bt_fill_smem(int TH, int NT) { return (size_t)2 * NT * 4 + (size_t)(TH + 1) * 4 + (size_t)((TH + 15) & ~15); }

__global__ void __launch_bounds__(256) long_tile_fill_kernel(const LongBtArgs a) {
    constexpr int CPT = LONG_BT_CPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int TH = a.TH, TW = a.TW, TWp = (TW + 31) & ~31;
    int* hbuf = reinterpret_cast<int*>(smem_raw);
    int* sleft = hbuf + 2 * NT;
    uint8_t* sq = reinterpret_cast<uint8_t*>(sleft + TH + 1);
    const int g = a.gap;
    const unsigned long long key = a.tiles[blockIdx.x];
    const long long tr = (long long)(key >> 32), tc = (long long)(key & 0xffffffffull);
    const long long r0 = tr * TH, c0 = tc * TW;
    const int h = (int)min((long long)TH, a.ie - r0), w = (int)min((long long)TW, a.je - c0);     // tiles on the end cell's row / column are cut there
    const int wt = (w + CPT - 1) / CPT;                                 // threads with at least one column
    uint32_t* __restrict__ slot = a.slots + (size_t)blockIdx.x * long_bt_slot_words(TH, TW);
    // ---- borders and sequence slices of the tile (row 0 and column 0 of the matrix are 0 and have no checkpoint) ------------
    for (int k = tid; k < h; k += NT) sq[k] = a.qry[r0 + k];
    for (int k = tid; k <= h; k += NT)
        sleft[k] = (tc > 0 && r0 + k > 0) ? a.colck[(tc - 1) * a.col_stride + r0 + k] : 0;            // H[r0 + k][c0]
    uint32_t rcv[CPT];
    int up[CPT];
    int diag0 = 0;                                                      // H[r - 1][column left of this thread's first]
    #pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = tid * CPT + k;
        rcv[k] = c < w ? (uint32_t)a.ref[c0 + c] : 0x100u;              // columns right of the cut never match; nobody reads them
        up[k] = (c < w && tr > 0) ? a.rowck[(tr - 1) * a.row_stride + c0 + 1 + c] : 0;              // H[r0][c0 + 1 + c]
    }
    if (tid < wt && tr > 0 && c0 + (long long)tid * CPT > 0) diag0 = a.rowck[(tr - 1) * a.row_stride + c0 + (long long)tid * CPT];
    __syncthreads();
    // ---- skewed sweep: at step s thread t fills tile row s - t of its CPT columns --------------------------------------------
    uint32_t acc[CPT];
    #pragma unroll
    for (int k = 0; k < CPT; ++k) acc[k] = 0;
    const int nsteps = h + wt - 1;
    for (int s = 0; s < nsteps; ++s) {
        const int r = s - tid;
        if (tid < wt && r >= 0 && r < h) {
            const int left_in = tid == 0 ? sleft[r + 1] : hbuf[((s - 1) & 1) * NT + tid - 1];
            const uint32_t qc = sq[r];
            int left = left_in, diag = diag0;
            const int sh = 2 * (r & 15);
            #pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int u = up[k];
                const int ug = u + g, lg = left + g;
                const int dg = diag + (qc == rcv[k] ? a.match : a.mismatch);
                const int v = __vimax3_s32_relu(ug, lg, dg);
                const uint32_t code = v == 0 ? C_STOP : (ug == v ? C_UP : (lg == v ? C_LEFT : C_DIAG));
                acc[k] |= code << sh;
                diag = u; up[k] = v; left = v;
            }
            diag0 = left_in;
            hbuf[(s & 1) * NT + tid] = left;
            if ((r & 15) == 15 || r == h - 1) {
                static_assert(CPT == 4, "one 16-byte store per thread");
                *reinterpret_cast<uint4*>(slot + (size_t)(r >> 4) * TWp + tid * CPT) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                #pragma unroll
                for (int k = 0; k < CPT; ++k) acc[k] = 0;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) long_walk_kernel(const LongBtArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int TH = a.TH, TW = a.TW, TWp = (TW + 31) & ~31;
    uint32_t* sdirs = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* sq = reinterpret_cast<uint8_t*>(sdirs + long_bt_slot_words(TH, TW));
    uint8_t* sr = sq + ((TH + 15) & ~15);
    __shared__ long long s_i, s_j, s_n, s_tiles;
    __shared__ int s_done, s_slot;
    if (tid == 0) { s_i = a.state[0]; s_j = a.state[1]; s_n = a.state[2]; s_done = (int)a.state[3]; s_tiles = a.state[4]; }
    __syncthreads();
    while (!s_done) {
        const long long i = s_i, j = s_j;
        const long long tr = (i - 1) / TH, tc = (j - 1) / TW;
        const long long r0 = tr * TH, c0 = tc * TW;
        const int h = (int)(i - r0), w = (int)(j - c0);                // the part of the tile at or above-left of the current cell
        // ---- which slot holds this tile?  none: the round is over ---------------------------------------------------------------
        if (tid == 0) s_slot = -1;
        __syncthreads();
        const unsigned long long key = ((unsigned long long)tr << 32) | (unsigned long long)tc;
        for (int b = tid; b < a.ntiles; b += NT) if (a.tiles[b] == key) s_slot = b;
        __syncthreads();
        const int b = s_slot;
        if (b < 0) break;
        // ---- stage the needed part of the slot and the two sequence slices ----------------------------------------------------------
        const uint32_t* __restrict__ slot = a.slots + (size_t)b * long_bt_slot_words(TH, TW);
        const int wr = (h + 15) >> 4, wq = (w + 3) >> 2;                // word rows, 16-byte groups per word row
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
            bool stopped = false;
            uint8_t* p0 = a.out + (a.cap - 1) - s_n, *p1 = p0 + a.cap, *p2 = p1 + a.cap;
            uint8_t* const p0_start = p0;
            while (ra > 0 && cb > 0) {
                const int x = ra - 1, y = cb - 1;
                const uint32_t d = (sdirs[(x >> 4) * TWp + y] >> (2 * (x & 15))) & 3u;
                if (d == C_STOP) { stopped = true; break; }
                const uint8_t qi = sq[x], rj = sr[y];
                const bool dg = d == C_DIAG, upm = d == C_UP;
                *p0-- = upm ? (uint8_t)'_' : rj;                                       // REF line: '_' where the query base has no partner
                *p1-- = dg ? (qi == rj ? (uint8_t)'*' : (uint8_t)'|') : (uint8_t)' ';
                *p2-- = (dg || upm) ? qi : (uint8_t)'_';
                ra -= (dg || upm) ? 1 : 0;
                cb -= (dg || !upm) ? 1 : 0;
            }
            const long long n = s_n + (long long)(p0_start - p0);
            s_n = n; s_i = r0 + ra; s_j = c0 + cb; s_tiles += 1;
            // on a tile edge the cell (s_i, s_j) belongs to the next tile, whose fill says through its STOP code whether H is 0 there
            s_done = stopped || s_i == 0 || s_j == 0;
        }
        __syncthreads();
    }
    if (tid == 0) { a.state[0] = s_i; a.state[1] = s_j; a.state[2] = s_n; a.state[3] = s_done; a.state[4] = s_tiles; }
}

}  // namespace dpx
