// host_long.cuh — host side of the long-pair path (csrc/longpair.cuh): kernel dispatch over (K, PACK, TABLE), the lane-width
// model, the single-GPU driver and the per-GPU stripe object of multi-GPU mode B.  Included by dpxalign.cu after the context,
// pool and CU() definitions; not a stand-alone header.
#pragma once

// ---- one long pair on one GPU: systolic array of warps over column blocks (longpair.cuh) ------------------------
struct LongPlan { int K; int capacity_warps; };

// mode bits: 1 = PACK (the travelling H and the query base share one 32-bit shuffle word; needs H < 2^23),
//            2 = TABLE (2-bit coded sequences, per-column score table; needs <= 4 symbols and int8 scores)
// mode bit 4 = CK: the checkpoint instantiation (row dumps compiled in; longtrace.cuh)
template <int K, bool PACK, bool TABLE, bool CK>
static int long_capacity(dpx_ctx* ctx, int* warps) {
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, long_sw_kernel<K, PACK, TABLE, CK>, 128, 0));
    *warps = per_sm * ctx->sm_count * 4;
    return DPX_OK;
}

template <int K, bool PACK, bool TABLE, bool CK>
static int long_launch(dpx_ctx* ctx, const LongArgs& a, cudaStream_t st) {
    const int blocks = (a.nwarps + 3) / 4;
    void* kargs[] = {(void*)&a};
    CU(cudaLaunchCooperativeKernel((void*)long_sw_kernel<K, PACK, TABLE, CK>, dim3(blocks), dim3(128), kargs, 0, st));
    return DPX_OK;
}

template <int K>
static int long_launch_m(dpx_ctx* ctx, int mode, const LongArgs& a, cudaStream_t st) {
    switch (mode & 7) {
        case 0: return long_launch<K, false, false, false>(ctx, a, st);
        case 1: return long_launch<K, true, false, false>(ctx, a, st);
        case 2: return long_launch<K, false, true, false>(ctx, a, st);
        case 3: return long_launch<K, true, true, false>(ctx, a, st);
        case 4: return long_launch<K, false, false, true>(ctx, a, st);
        case 5: return long_launch<K, true, false, true>(ctx, a, st);
        case 6: return long_launch<K, false, true, true>(ctx, a, st);
        default: return long_launch<K, true, true, true>(ctx, a, st);
    }
}
template <int K>
static int long_capacity_m(dpx_ctx* ctx, int mode, int* warps) {
    switch (mode & 7) {
        case 0: return long_capacity<K, false, false, false>(ctx, warps);
        case 1: return long_capacity<K, true, false, false>(ctx, warps);
        case 2: return long_capacity<K, false, true, false>(ctx, warps);
        case 3: return long_capacity<K, true, true, false>(ctx, warps);
        case 4: return long_capacity<K, false, false, true>(ctx, warps);
        case 5: return long_capacity<K, true, false, true>(ctx, warps);
        case 6: return long_capacity<K, false, true, true>(ctx, warps);
        default: return long_capacity<K, true, true, true>(ctx, warps);
    }
}

// K = 32 exists for the score-table kernels only (mode bit 2); the byte-compare kernels stop at 16 columns per lane
template <int K>
static int long_launch_t(dpx_ctx* ctx, int mode, const LongArgs& a, cudaStream_t st) {
    switch (mode & 7) {
        case 2: return long_launch<K, false, true, false>(ctx, a, st);
        case 3: return long_launch<K, true, true, false>(ctx, a, st);
        case 6: return long_launch<K, false, true, true>(ctx, a, st);
        case 7: return long_launch<K, true, true, true>(ctx, a, st);
    }
    ctx->err = "32 columns per lane need the score-table kernel"; return DPX_ERR_INVALID;
}
template <int K>
static int long_capacity_t(dpx_ctx* ctx, int mode, int* warps) {
    switch (mode & 7) {
        case 2: return long_capacity<K, false, true, false>(ctx, warps);
        case 3: return long_capacity<K, true, true, false>(ctx, warps);
        case 6: return long_capacity<K, false, true, true>(ctx, warps);
        case 7: return long_capacity<K, true, true, true>(ctx, warps);
    }
    ctx->err = "32 columns per lane need the score-table kernel"; return DPX_ERR_INVALID;
}

static int long_launch_k(dpx_ctx* ctx, int K, int mode, const LongArgs& a, cudaStream_t st) {
    switch (K) {
        case 2: return long_launch_m<2>(ctx, mode, a, st);
        case 4: return long_launch_m<4>(ctx, mode, a, st);
        case 8: return long_launch_m<8>(ctx, mode, a, st);
        case 32: return long_launch_t<32>(ctx, mode, a, st);
        default: return long_launch_m<16>(ctx, mode, a, st);
    }
}

static int long_capacity_k(dpx_ctx* ctx, int K, int mode, int* warps) {
    switch (K) {
        case 2: return long_capacity_m<2>(ctx, mode, warps);
        case 4: return long_capacity_m<4>(ctx, mode, warps);
        case 8: return long_capacity_m<8>(ctx, mode, warps);
        case 32: return long_capacity_t<32>(ctx, mode, warps);
        default: return long_capacity_m<16>(ctx, mode, warps);
    }
}

// Byte -> 2-bit code map over both sequences of a long pair; returns the number of distinct symbols (codes are only
// meaningful when it is <= 4).  Host pass over a few MB, off the kernel's path.
static int long_alphabet(const char* ref, size_t R, const char* qry, size_t Q, uint8_t code[256]) {
    bool present[256] = {false};
    for (size_t i = 0; i < R; ++i) present[(uint8_t)ref[i]] = true;
    for (size_t i = 0; i < Q; ++i) present[(uint8_t)qry[i]] = true;
    int n = 0;
    for (int c = 0; c < 256; ++c) { code[c] = 0; if (present[c]) code[c] = (uint8_t)(n++ & 3); }
    return n;
}

static bool long_table_ok(const dpx_params* p, size_t R, size_t Q) {
    const int ms = p->match - p->gap_open, xs = p->mismatch - p->gap_open;
    // int8 table entries; the row-maximum keys are h*16 + column, so h must stay below 2^27
    return ms >= -128 && ms <= 127 && xs >= -128 && xs <= 127 && p->match > 0 &&
           (long double)p->match * (long double)std::min(R, Q) < 1.3e8L;
}

static bool long_can_pack(const dpx_params* p, size_t R, size_t Q) {
    return (long double)p->match * (long double)std::min(R, Q) < 8.0e6L && p->match > 0;
}

static long long pow2_at_least(long long v) { long long p = 64; while (p < v) p <<= 1; return p; }

// K: columns per lane.  Measured on B200 (tools/long_sweep.py): with at most one warp per SM sub-partition a row step costs
// L(K) = 172 / 296 / 344 cycles for K = 8 / 16 / 32 (a dependent chain of 2 DPX ops per cell plus the shuffle / ring overhead of
// the step); with w warps per sub-partition it stretches by 1 + 0.64 (w - 1) (1 + 1.43 (w - 1) at K = 32, whose 127 registers
// leave less room to overlap).  The chain advances one row per step, so take the K that minimises the step time; ties go to the
// wider lane (fewer warps = shorter pipeline fill).  Narrower lanes (K = 4, 2) only pay for references of a few thousand bases,
// where they are what spreads the work over more than a handful of warps.
static int long_pick_k(dpx_ctx* ctx, long long R_local, bool allow32 = false) {
    int best_k = 16; double best = 1e300;
    if (R_local < 4096) return R_local < 1024 ? 2 : 4;
    for (int K : {32, 16, 8}) {
        if (K == 32 && !allow32) continue;            // 32 columns per lane: score-table kernels only, keys h * 32 + column must fit int32
        const double L = K == 32 ? 344.0 : K == 16 ? 296.0 : 172.0;
        const double w = (double)((R_local + 32LL * K - 1) / (32LL * K)) / (4.0 * ctx->sm_count);
        const double cost = L * (w <= 1.0 ? 1.0 : 1.0 + (K == 32 ? 1.43 : 0.64) * (w - 1.0));
        if (cost < best * 0.999) { best = cost; best_k = K; }
    }
    return best_k;
}

// Checkpoint mode of the forward pass (longtrace.cuh): every warp also stores the column it reads from its left neighbour, so
// the left edge of every column block, H[i][c * CW] for all rows i, is still there when the kernel has finished; and every lane
// stores its columns of every TH-th row.  Together: the top row and the left column of every TH x CW tile of the matrix.
struct LongCkpt {
    int32_t* base = nullptr;              // H[i][(c + 1) * CW] at base[c * stride + i], i = 1..Q
    long long stride = 0;                 // entries per column boundary (>= Q + 1)
    long long n = 0;                      // column boundaries = column blocks - 1
    int CW = 0;                           // columns per block = tile width
    int32_t* rbase = nullptr;             // H[(m + 1) * TH][j] at rbase[m * rstride + j - 1], j = 1..R
    long long rstride = 0, rn = 0;        // entries per checkpoint row, checkpoint rows = floor(Q / TH)
    int TH = 0;                           // tile height (a power of two)
    void free() { if (base) cudaFree(base); if (rbase) cudaFree(rbase); base = rbase = nullptr; }
};

static int long_pair_single(dpx_ctx* ctx, const dpx_params* p, const char* ref, size_t R, const char* qry, size_t Q,
                            int32_t* score, int64_t* end_row, int64_t* end_col, LongCkpt* ck = nullptr) {
    cudaStream_t st = ctx->stream;
    uint8_t code[256];
    const bool table = long_table_ok(p, R, Q) && long_alphabet(ref, R, qry, Q, code) <= 4 && !ctx->opt.long_notable;
    const bool allow32 = table && (long double)p->match * (long double)std::min(R, Q) < 6.0e7L;
    int K = long_pick_k(ctx, (long long)R, allow32);
    if (const int k = ctx->opt.long_k) { if (k != 32 || allow32) K = k; }                                                    // dpx_set_option("long_k")
    int capacity = 0;
    const int mode = (long_can_pack(p, R, Q) ? 1 : 0) | (table ? 2 : 0) | (ck ? 4 : 0);
    { int s = long_capacity_k(ctx, K, mode, &capacity); if (s) return s; }
    if (const int c = ctx->opt.long_cap) { if (c >= 4 && c < capacity) capacity = c & ~3; }                                   // dpx_set_option("long_cap"): force passes
    if (capacity < 4) { ctx->err = "long-pair kernel does not fit"; return DPX_ERR_RANGE; }
    const long long CW = 32LL * K;
    const long long nw_total = ((long long)R + CW - 1) / CW;
    const long long passes = (nw_total + capacity - 1) / capacity;
    const long long nw_pass = (nw_total + passes - 1) / passes;
    const long long RING = 2048;

    uint8_t *d_ref = nullptr, *d_qry = nullptr;
    unsigned long long *d_rings = nullptr, *d_full[2] = {nullptr, nullptr}; int32_t* d_bs = nullptr;
    long long *d_cnt = nullptr, *d_br = nullptr, *d_bc = nullptr;
    LongChan* d_chans = nullptr; int* d_err = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        DevPool& P = ctx->pool;
        P.release(d_ref); P.release(d_qry); P.release(d_rings); P.release(d_full[0]); P.release(d_full[1]); P.release(d_bs);
        P.release(d_cnt); P.release(d_br); P.release(d_bc); P.release(d_chans); P.release(d_err);
    };
#define LCU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return DPX_ERR_CUDA; } } while (0)
    bool ok = pool_alloc(ctx, &d_ref, R + 16) && pool_alloc(ctx, &d_qry, Q + 16) && pool_alloc(ctx, &d_rings, (size_t)(nw_pass * RING)) &&
              pool_alloc(ctx, &d_cnt, (size_t)(2 * (nw_pass + 2))) && pool_alloc(ctx, &d_bs, (size_t)nw_pass) &&
              pool_alloc(ctx, &d_br, (size_t)nw_pass) && pool_alloc(ctx, &d_bc, (size_t)nw_pass) &&
              pool_alloc(ctx, &d_chans, (size_t)(nw_pass + 1)) && pool_alloc(ctx, &d_err, 1);
    const long long FULLSZ = pow2_at_least((long long)Q + 2);      // ring sizes are powers of two; this one never wraps
    if (ok && passes > 1) ok = pool_alloc(ctx, &d_full[0], (size_t)FULLSZ) && pool_alloc(ctx, &d_full[1], (size_t)FULLSZ);
    if (!ok) { cleanup(); return DPX_ERR_NOMEM; }
    if (ck) {
        ck->CW = (int)CW; ck->n = nw_total - 1; ck->stride = ((long long)Q + 1 + 31) & ~31LL; ck->base = nullptr;
        ck->TH = (int)std::min<long long>(CW, 512); ck->rn = (long long)Q / ck->TH; ck->rstride = ((long long)R + 3) & ~3LL; ck->rbase = nullptr;
        const size_t bytes = sizeof(int32_t) * (size_t)ck->n * (size_t)ck->stride, rbytes = sizeof(int32_t) * (size_t)ck->rn * (size_t)ck->rstride;
        if ((bytes && cudaMalloc(&ck->base, bytes) != cudaSuccess) || (rbytes && cudaMalloc(&ck->rbase, rbytes) != cudaSuccess)) {
            cudaGetLastError(); cleanup();
            ctx->err = "long-pair traceback: " + std::to_string((bytes + rbytes) >> 20) + " MiB of checkpoints do not fit in device memory";
            return DPX_ERR_NOMEM;
        }
    }
    std::vector<uint8_t> cref, cqry;
    if (table) {
        cref.resize(R); cqry.resize(Q);
        for (size_t i = 0; i < R; ++i) cref[i] = code[(uint8_t)ref[i]];
        for (size_t i = 0; i < Q; ++i) cqry[i] = code[(uint8_t)qry[i]];
    }
    LCU(cudaMemcpyAsync(d_ref, table ? (const char*)cref.data() : ref, R, cudaMemcpyHostToDevice, st));
    LCU(cudaMemcpyAsync(d_qry, table ? (const char*)cqry.data() : qry, Q, cudaMemcpyHostToDevice, st));
    LCU(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    LCU(cudaStreamSynchronize(st));

    int32_t best = 0; long long brow = 0, bcol = 0;
    std::vector<LongChan> chans((size_t)nw_pass + 1);
    std::vector<int32_t> h_bs((size_t)nw_pass); std::vector<long long> h_br((size_t)nw_pass), h_bc((size_t)nw_pass);
    for (long long ps = 0; ps < passes; ++ps) {
        const long long w0 = ps * nw_pass, nw = std::min(nw_pass, nw_total - w0);
        if (nw <= 0) break;
        // credit counters and ring tags start from zero (a tag of 0 never equals a row >= 1)
        LCU(cudaMemsetAsync(d_cnt, 0, sizeof(long long) * (size_t)(2 * (nw_pass + 2)), st));
        LCU(cudaMemsetAsync(d_rings, 0, sizeof(unsigned long long) * (size_t)(nw_pass * RING), st));
        if (ps + 1 < passes) LCU(cudaMemsetAsync(d_full[ps & 1], 0, sizeof(unsigned long long) * (size_t)FULLSZ, st));
        long long* cred = d_cnt;
        for (long long c = 0; c <= nw; ++c) {
            LongChan ch{};
            if (c == 0) {
                if (ps > 0) { ch.ring = d_full[(ps - 1) & 1]; ch.size = FULLSZ; ch.credit = nullptr; }
            } else if (c == nw) {
                if (ps + 1 < passes) { ch.ring = d_full[ps & 1]; ch.size = FULLSZ; ch.credit = nullptr; }
            } else {
                ch.ring = d_rings + (c - 1) * RING; ch.size = RING; ch.credit = cred + c;
            }
            chans[(size_t)c] = ch;
        }
        LCU(cudaMemcpyAsync(d_chans, chans.data(), sizeof(LongChan) * (size_t)(nw + 1), cudaMemcpyHostToDevice, st));
        LongArgs a{};
        a.ref = d_ref; a.qry = d_qry; a.Q = (long long)Q; a.R_local = (long long)R; a.col0 = w0 * CW; a.col_offset = 0;
        a.match = p->match; a.mismatch = p->mismatch; a.gap = p->gap_open; a.nwarps = (int)nw; a.chans = d_chans;
        a.best_score = d_bs; a.best_row = d_br; a.best_col = d_bc; a.error_flag = d_err; a.system_scope = 0;
        a.tab_match = p->match - p->gap_open; a.tab_mismatch = p->mismatch - p->gap_open; a.sixteen = 16u;
        if (ck) {
            a.ck_base = ck->base; a.ck_stride = ck->stride; a.ck_first = w0;
            a.rk_base = ck->rbase; a.rk_stride = ck->rstride; a.rk_shift = 0;
            while ((1 << a.rk_shift) < ck->TH) ++a.rk_shift;
        }
        { int s = long_launch_k(ctx, K, mode, a, st); if (s) { cleanup(); return s; } }
        LCU(cudaMemcpyAsync(h_bs.data(), d_bs, sizeof(int32_t) * (size_t)nw, cudaMemcpyDeviceToHost, st));
        LCU(cudaMemcpyAsync(h_br.data(), d_br, sizeof(long long) * (size_t)nw, cudaMemcpyDeviceToHost, st));
        LCU(cudaMemcpyAsync(h_bc.data(), d_bc, sizeof(long long) * (size_t)nw, cudaMemcpyDeviceToHost, st));
        int err = 0;
        LCU(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        if (err) { ctx->err = "long-pair pipeline watchdog fired"; cleanup(); return DPX_ERR_CUDA; }
        for (long long w = 0; w < nw; ++w) {
            const int32_t s = h_bs[(size_t)w]; const long long r = h_br[(size_t)w], c = h_bc[(size_t)w];
            if (s > best || (s == best && s > 0 && (r < brow || (r == brow && c < bcol)))) { best = s; brow = r; bcol = c; }
        }
    }
#undef LCU
    cleanup();
    *score = best; if (end_row) *end_row = brow; if (end_col) *end_col = bcol;
    return DPX_OK;
}

// ---- alignment strings of one long pair: checkpointed forward passes + tile-by-tile walk (longtrace.cuh) ----------------
// lines: library-allocated (dpx_free) blob of three NUL-terminated lines REF, REL, QRY, each `len` characters, at offsets
// 0, len + 1, 2 (len + 1) — the three lines LinearSmithWaterman::print_results writes (c++/LinearSmithWaterman.cpp:259-285).
struct LongTraceStats { double fwd_ms = 0, walk_ms = 0; long long tiles = 0, rounds = 0; int TH = 0, TW = 0; };

static int long_pair_strings(dpx_ctx* ctx, const dpx_params* p, const char* ref, size_t R, const char* qry, size_t Q,
                             int32_t* score, int64_t* end_row, int64_t* end_col, int64_t* start_row, int64_t* start_col,
                             char** lines, size_t* len, LongTraceStats* stats) {
    cudaStream_t st = ctx->stream;
    LongCkpt colck;
    uint8_t *d_ref = nullptr, *d_qry = nullptr, *d_out = nullptr; long long* d_res = nullptr;
    unsigned long long* d_tiles = nullptr; uint32_t *d_slots = nullptr, *d_edges = nullptr; int4* d_segs = nullptr; long long* d_seg_off = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        colck.free();
        ctx->pool.release(d_ref); ctx->pool.release(d_qry); ctx->pool.release(d_out); ctx->pool.release(d_res); ctx->pool.release(d_tiles); ctx->pool.release(d_slots); ctx->pool.release(d_edges); ctx->pool.release(d_segs); ctx->pool.release(d_seg_off);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    };
#define TCU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return DPX_ERR_CUDA; } } while (0)
    TCU(cudaEventCreate(&ev[0])); TCU(cudaEventCreate(&ev[1]));
    auto timed = [&](double* ms, auto&& fn) -> int {
        cudaEventRecord(ev[0], st);
        const int r = fn();
        cudaEventRecord(ev[1], st); cudaEventSynchronize(ev[1]);
        float f = 0; cudaEventElapsedTime(&f, ev[0], ev[1]); *ms = f;
        return r;
    };
    LongTraceStats ls;
    int32_t sc = 0; int64_t ie = 0, je = 0;
    // forward pass with checkpoints: H on every TW-th column and every TH-th row
    LongCkpt ck;
    { int r = timed(&ls.fwd_ms, [&] { return long_pair_single(ctx, p, ref, R, qry, Q, &sc, &ie, &je, &ck); }); if (r) { ck.free(); cleanup(); return r; } }
    colck = ck;                                                  // owned by cleanup() from here on
    *score = sc; if (end_row) *end_row = ie; if (end_col) *end_col = je;
    if (start_row) *start_row = ie; if (start_col) *start_col = je;
    *len = 0;
    if (sc <= 0) {                                               // nothing aligns: three empty lines (:253-257)
        char* blob = (char*)g_host.take(4);
        if (!blob) { cleanup(); return DPX_ERR_NOMEM; }
        blob[0] = blob[1] = blob[2] = 0; *lines = blob;
        cleanup();
        if (stats) *stats = ls;
        return DPX_OK;
    }
    const int TW = colck.CW, TH = colck.TH;
    const long long cap = (long long)ie + (long long)je;
    if (!pool_alloc(ctx, &d_ref, (size_t)je + 16) || !pool_alloc(ctx, &d_qry, (size_t)ie + 16) || !pool_alloc(ctx, &d_out, (size_t)(3 * cap) + 16) ||
        !pool_alloc(ctx, &d_res, 8)) { cleanup(); return DPX_ERR_NOMEM; }
    TCU(cudaMemcpyAsync(d_ref, ref, (size_t)je, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_qry, qry, (size_t)ie, cudaMemcpyHostToDevice, st));
    LongBtArgs a{};
    a.ref = d_ref; a.qry = d_qry; a.ie = ie; a.je = je; a.match = p->match; a.mismatch = p->mismatch; a.gap = p->gap_open;
    a.TH = TH; a.TW = TW; a.colck = colck.base; a.col_stride = colck.stride;
    a.rowck = colck.rbase - 1; a.row_stride = colck.rstride;      // 0-based columns: rowck[r * stride + j] is column j >= 1
    a.out = d_out; a.cap = cap; a.state = d_res;
    // ---- rounds: predict the tiles ahead of the walk, fill them all at once, walk until the walk stops or leaves them ----
    const int nt = std::max(32, ((TW + LONG_BT_CPT - 1) / LONG_BT_CPT + 31) & ~31);
    const size_t fill_smem = long_bt_fill_smem(TH, nt), walk_smem = long_bt_walk_smem(TH, TW), slot_words = long_bt_slot_words(TH, TW);
    if (nt > 256 || walk_smem > (size_t)200 * 1024) { ctx->err = "long-pair traceback: tile does not fit shared memory"; cleanup(); return DPX_ERR_RANGE; }
    int max_tiles = 2 * ctx->sm_count;
    if (ctx->opt.long_bt_tiles >= 1) max_tiles = ctx->opt.long_bt_tiles;                                                      // dpx_set_option("long_bt_tiles"): short rounds
    if (!pool_alloc(ctx, &d_tiles, (size_t)max_tiles) || !pool_alloc(ctx, &d_slots, slot_words * (size_t)max_tiles) ||
        !pool_alloc(ctx, &d_edges, long_bt_edge_words(TH, TW) * (size_t)max_tiles) || !pool_alloc(ctx, &d_segs, (size_t)max_tiles) ||
        !pool_alloc(ctx, &d_seg_off, (size_t)max_tiles)) { cleanup(); return DPX_ERR_NOMEM; }
    TCU(max_dyn_smem(ctx, long_emit_kernel));
    TCU(max_dyn_smem(ctx, long_tile_fill_kernel));
    a.slots = d_slots; a.tiles = d_tiles; a.edges = d_edges; a.segs = d_segs; a.seg_off = d_seg_off;
    long long state[8] = {ie, je, 0, 0, 0, 0, 0, 0};              // row, column, characters written, done, tiles walked, segments, error
    TCU(cudaMemcpyAsync(d_res, state, sizeof(state), cudaMemcpyHostToDevice, st));
    double di = 1.0, dj = 1.0;                                    // the walk's direction, rows and columns per step of the prediction line
    std::vector<unsigned long long> tiles;
    const long long margin = std::max<long long>(TW / 8, 16);
    cudaEventRecord(ev[0], st);
    while (!state[3]) {
        // the band: tiles under the line (i - t di, j - t dj) and `margin` columns either side of it, the current tile first
        tiles.clear();
        const long long i0 = state[0], j0 = state[1];
        const double step = (double)std::min(TH, TW) / 8.0;
        for (double t = 0; (int)tiles.size() < max_tiles; t += step) {
            const long long ii = i0 - (long long)(t * di), jj = j0 - (long long)(t * dj);
            if (ii < 1 || jj < 1) break;
            for (long long dlt : {0LL, -margin, margin}) {
                const long long jc = std::min<long long>(std::max<long long>(jj + dlt, 1), je);
                const unsigned long long key = ((unsigned long long)((ii - 1) / TH) << 32) | (unsigned long long)((jc - 1) / TW);
                if (std::find(tiles.begin(), tiles.end(), key) == tiles.end() && (int)tiles.size() < max_tiles) tiles.push_back(key);
            }
        }
        a.ntiles = (int)tiles.size();
        TCU(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(unsigned long long) * tiles.size(), cudaMemcpyHostToDevice, st));
        long_tile_fill_kernel<<<a.ntiles, nt, fill_smem, st>>>(a);
        long_chain_kernel<<<1, 32, 0, st>>>(a);
        long_emit_kernel<<<a.ntiles, 256, walk_smem, st>>>(a);
        TCU(cudaGetLastError());
        TCU(cudaMemcpyAsync(state, d_res, sizeof(state), cudaMemcpyDeviceToHost, st));
        TCU(cudaStreamSynchronize(st));
        ++ls.rounds;
        const long long mi = i0 - state[0], mj = j0 - state[1];
        if (mi + mj >= 64) { const double m = (double)std::max(mi, mj); di = mi / m; dj = mj / m; }
        else { di = dj = 1.0; }
        if (state[6]) { ctx->err = "long-pair traceback: transfer tables and directions disagree"; cleanup(); return DPX_ERR_CUDA; }
        if (!state[3] && mi == 0 && mj == 0) { ctx->err = "long-pair traceback: the walk made no progress"; cleanup(); return DPX_ERR_CUDA; }
    }
    cudaEventRecord(ev[1], st); cudaEventSynchronize(ev[1]);
    { float f = 0; cudaEventElapsedTime(&f, ev[0], ev[1]); ls.walk_ms = f; }
    const long long L = state[2];
    const long long res[3] = {L, state[0], state[1]};
    ls.tiles = state[4]; ls.TH = TH; ls.TW = TW;
    if (L <= 0 || L > cap) { ctx->err = "long-pair traceback: walk returned an impossible length"; cleanup(); return DPX_ERR_CUDA; }
    char* blob = (char*)g_host.take((size_t)(3 * (L + 1)));
    if (!blob) { cleanup(); return DPX_ERR_NOMEM; }
    for (int k = 0; k < 3; ++k) {
        if (cudaMemcpyAsync(blob + (size_t)k * (size_t)(L + 1), d_out + (size_t)k * (size_t)cap + (size_t)(cap - L), (size_t)L, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            dpx_free(blob); ctx->err = "long-pair traceback: download failed"; cleanup(); return DPX_ERR_CUDA;
        }
        blob[(size_t)k * (size_t)(L + 1) + (size_t)L] = 0;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { dpx_free(blob); ctx->err = "long-pair traceback: download failed"; cleanup(); return DPX_ERR_CUDA; }
#undef TCU
    *lines = blob; *len = (size_t)L;
    if (start_row) *start_row = res[1]; if (start_col) *start_col = res[2];
    if (stats) *stats = ls;
    cleanup();
    return DPX_OK;
}

// ---- multi-GPU mode B: one column stripe of a long pair per GPU -------------------------------------------------
struct dpx_stripe {
    dpx_ctx* ctx = nullptr;
    dpx_params params{};
    size_t R_local = 0, col_offset = 0, Q = 0;
    int index = 0, n = 1, K = 8, nw = 0; int mode = 0;
    static constexpr long long XRING = 65536, RING = 2048;
    // exchange buffer (own memory, exported over CUDA IPC): [1] out credit (written by the next stripe), inbox ring of
    // tagged 8-byte entries at byte 128 (written by the previous stripe)
    char* xbuf = nullptr;
    char* prev_x = nullptr; char* next_x = nullptr;          // neighbours' exchange buffers (peer mappings)
    uint8_t *d_ref = nullptr, *d_qry = nullptr;
    unsigned long long* d_rings = nullptr; int32_t* d_bs = nullptr;
    long long *d_cnt = nullptr, *d_br = nullptr, *d_bc = nullptr;
    LongChan* d_chans = nullptr; int* d_err = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool launched = false;
};

