// host_run.cuh — kernel selection and launch of one batch run (dpx_batch_run): eligibility + constant plans of the specialised
// kernels (shortread.cuh, pairwf.cuh, band.cuh), the chunked traceback pipeline, and the general wavefront fall-back.
// Included by dpxalign.cu after the context, pool, batch and CU() definitions; not a stand-alone header.
#pragma once

template <int ALGO, bool TB, int K>
static int query_wf(dpx_ctx* ctx, int slots_wanted, int* blocks_out) {
    int per_sm = 0;
    CU(occupancy(ctx, &per_sm, wf_fill_kernel<ALGO, TB, K>, 128, 0));
    if (per_sm < 1) per_sm = 1;
    int blocks = std::min(ctx->sm_count * per_sm, (slots_wanted + 3) / 4);
    *blocks_out = std::max(blocks, 1);
    return DPX_OK;
}

template <int ALGO, bool TB>
static int dispatch_wf_k(dpx_ctx* ctx, cudaStream_t st, int K, const WfArgs& a, int blocks, bool query_only, int slots, int* blocks_out) {
    if (K == 4) { if (query_only) return query_wf<ALGO, TB, 4>(ctx, slots, blocks_out); wf_fill_kernel<ALGO, TB, 4><<<blocks, 128, 0, st>>>(a); }
    else        { if (query_only) return query_wf<ALGO, TB, 8>(ctx, slots, blocks_out); wf_fill_kernel<ALGO, TB, 8><<<blocks, 128, 0, st>>>(a); }
    return DPX_OK;
}

static int dispatch_wf(dpx_ctx* ctx, cudaStream_t st, int algo, bool tb, int K, const WfArgs& a, int blocks, bool query_only, int slots, int* blocks_out) {
    switch (algo) {
        case DPX_ALGO_LNW: return tb ? dispatch_wf_k<DPX_ALGO_LNW, true>(ctx, st, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_LNW, false>(ctx, st, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_ANW: return tb ? dispatch_wf_k<DPX_ALGO_ANW, true>(ctx, st, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_ANW, false>(ctx, st, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_LSW: return tb ? dispatch_wf_k<DPX_ALGO_LSW, true>(ctx, st, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_LSW, false>(ctx, st, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_BSW: return tb ? dispatch_wf_k<DPX_ALGO_BSW, true>(ctx, st, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_BSW, false>(ctx, st, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_ABSW: return tb ? dispatch_wf_k<DPX_ALGO_ABSW, true>(ctx, st, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_ABSW, false>(ctx, st, K, a, blocks, query_only, slots, blocks_out);
    }
    return DPX_ERR_INVALID;
}

template <int G, int K>
static int run_short(dpx_ctx* ctx, dpx_batch* b, SrArgs a, bool track, bool wide) {
    const int gpb = 128 / G;
    a.bnd_stride = b->max_r + 2 * G + 4;                       // G words of slack in front (shortread.cuh: bnd)
    a.rsel_stride = (b->max_r + 2 * G + 18 + 1) & ~1;          // the table is written 16 entries (one packed word) at a time
    const size_t smem = (size_t)gpb * ((size_t)a.bnd_stride * 4 + (size_t)a.rsel_stride * 2);
    auto launch = [&](auto kern) -> int {
        CU(max_dyn_smem(ctx, kern));
        int per_sm = 0;
        CU(occupancy(ctx, &per_sm, kern, 128, smem));
        if (per_sm < 1) { ctx->err = "short-read kernel does not fit on an SM"; return DPX_ERR_RANGE; }
        const int warps_needed = (a.n_slots + (32 / G) - 1) / (32 / G);   // (n_slots: pair duos, or single pairs in WIDE mode)
        int blocks = std::min(ctx->sm_count * per_sm, (warps_needed + 3) / 4);
        if (blocks < 1) blocks = 1;
        kern<<<blocks, 128, smem, b->stream>>>(a);
        CU(cudaGetLastError());
        return DPX_OK;
    };
    if (track) return wide ? launch(sr_lsw_kernel<G, K, true, true>) : launch(sr_lsw_kernel<G, K, true, false>);
    return wide ? launch(sr_lsw_kernel<G, K, false, true>) : launch(sr_lsw_kernel<G, K, false, false>);
}

// Eligibility of the packed int16x2 short-read kernel (shortread.cuh).
static bool short_eligible(const dpx_batch* b, const dpx_params* p, int* B_out, bool* wide, int* kbits_out) {
    // <= 4 symbols: the 2-bit packed batch, two pairs per lane group; 5..8 symbols (byte codes present): one pair per group (WIDE)
    const bool w = !b->packed2 && b->d_codes != nullptr;
    if (p->algo != DPX_ALGO_LSW || (p->flags & DPX_OUT_STRINGS) || (!b->packed2 && !w) || b->ctx->opt.no_shortread) return false;
    const int m = p->match, x = p->mismatch, g = p->gap_open;
    if (!(m > 0 && x < 0 && g < 0)) return false;                 // pads must stay strictly below real cells
    if (m - g > 127 || x - g < -128 || x - g > 127 || g < -4096) return false;
    if (b->max_r > 4096 || b->max_q > 65535) return false;
    // 16 pair groups per block, each with a boundary row (4 B per column) and a column table (2 B): must fit one block's shared memory
    if ((size_t)16 * ((size_t)(b->max_r + 20) * 4 + (size_t)(b->max_r + 36) * 2) > (size_t)227 * 1024) return false;
    const int B = std::max(2, -g);
    // position bits: (Hmax + B) << k < 65536 (unsigned 16-bit keys); one of the k bits marks the upper row of a row pair, the other
    // k-1 count steps inside blocks of 2^(k-1) steps; at most 256 blocks per pass
    const long long top = (long long)m * std::min(b->max_r, b->max_q) + B;
    int k = 0;
    while (k < 8 && (top << (k + 1)) < 65536) ++k;
    if (k < 3) return false;                  // below that the fold every 2^(k-1) steps costs more than it saves
    if (((long long)b->max_r + 16) >> (k - 1) >= 255) return false;
    *B_out = B; *wide = w; *kbits_out = k;
    return true;
}

// Scratch of the per-batch CUB passes stays with the batch until it is released (batch_release runs after the stream has
// drained): no host sync here, so the chunks of the one-call pipeline keep streaming when the lengths are ragged.
static int ensure_order(dpx_batch* b) {
    dpx_ctx* ctx = b->ctx;
    if (b->uniform || b->d_order) return DPX_OK;
    const int n = (int)b->n_pairs;
    unsigned long long *k_in = nullptr, *k_out = nullptr; int32_t* v_in = nullptr;
    if (!pool_alloc(ctx, &k_in, n) || !pool_alloc(ctx, &k_out, n) || !pool_alloc(ctx, &v_in, n)) { ctx->pool.release(k_in); ctx->pool.release(k_out); ctx->pool.release(v_in); return DPX_ERR_NOMEM; }
    b->d_scratch.push_back(k_in); b->d_scratch.push_back(k_out); b->d_scratch.push_back(v_in);
    if (!pool_alloc(ctx, &b->d_order, n)) return DPX_ERR_NOMEM;
    sched_keys_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(b->d_pairs, n, k_in, v_in);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, b->d_order, n, 0, 64, b->stream);
    void* tmp = ctx->pool.alloc(tmp_bytes);
    if (!tmp) return DPX_ERR_NOMEM;
    b->d_scratch.push_back(tmp);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, b->d_order, n, 0, 64, b->stream));
    return DPX_OK;
}

static int ensure_str_off(dpx_batch* b) {
    dpx_ctx* ctx = b->ctx;
    if (b->d_str_off) return DPX_OK;
    const int n = (int)b->n_pairs;
    if (!pool_alloc(ctx, &b->d_str_off, n + 1) || !pool_alloc(ctx, &b->d_str_start, n)) return DPX_ERR_NOMEM;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b->d_str_len, b->d_str_off, n, b->stream);
    void* tmp = ctx->pool.alloc(tmp_bytes);
    if (!tmp) return DPX_ERR_NOMEM;
    b->d_scratch.push_back(tmp);
    CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, b->d_str_len, b->d_str_off, n, b->stream));
    if (!pool_alloc(ctx, &b->d_strings, (size_t)b->info.str_bytes + 1)) return DPX_ERR_NOMEM;
    return DPX_OK;
}


// ---- packed two-pair Needleman-Wunsch path (pairwf.cuh): plan = every constant of the 4*X + BIAS + code arithmetic ----
struct PwPlan { int K; bool packed, wide; uint32_t lut_lo, lut_hi, ext2, addc, addc3, zero2; int b0, b1, bstep, dec_sub, dec_add; };

static bool pairwf_eligible(const dpx_batch* b, const dpx_params* p, PwPlan* pl) {
    const bool sw = p->algo == DPX_ALGO_LSW;
    const bool wide = !b->packed2 && b->d_codes != nullptr;       // 5..8 symbols: int32, both table registers for one pair
    if ((p->algo != DPX_ALGO_LNW && p->algo != DPX_ALGO_ANW && !sw) || (!b->packed2 && !wide) || b->ctx->opt.no_pairwf) return false;
    if (sw && !(p->flags & DPX_OUT_STRINGS)) return false;      // score / end cell alone: the short-read kernel (or the int32 wavefront)
    const bool aff = p->algo == DPX_ALGO_ANW;
    const long long m = p->match, x = p->mismatch, go = p->gap_open, ge = aff ? p->gap_extend : 0;
    const long long open = aff ? go + ge : go;                   // cost of the first gap column: "goe" (Gotoh) or g (linear)
    if (open >= 0 || ge > 0 || go > 0) return false;             // the add constant must be negative (always-carry rule)
    if (sw && !(m > 0 && x < 0)) return false;                   // pads must stay strictly below the maximum
    const long long code = aff ? 3 : 1;
    const long long tm = 4 * (m - open) - code, tx = 4 * (x - open) - code;
    if (tm < 0 || tm > 127 || tx < 0 || tx > 127) return false;     // table bytes are sign-extended by the selector
    if (b->max_r > 12000) return false;                          // the per-warp column table (2 B per column, 4 warps per block) must leave 2 blocks per SM
    // rows per lane: 8; packed Gotoh WITHOUT traceback on queries of more than one 256-row pass takes 16 (the per-step work that does
    // not depend on the rows -- shuffles, selector load, predicates: a fifth of the step at 8 rows -- is paid half as often).  With
    // traceback the 16-row kernel needs 165 registers: three blocks per SM fill the register file, the walk of the previous chunk
    // finds no room beside the fill and the chunk pipeline of section 5 stalls (measured: fill alone 34.0 ms instead of 35.4, the
    // pipelined step 46.4 instead of 36.5), so that path stays at 8.
    int K = (aff && !wide && b->max_q > 256 && !(p->flags & DPX_OUT_STRINGS) && !b->ctx->opt.pairwf_k8) ? 16 : 8;
    bool packed = false; long long B = 0;
    for (;;) {
        const long long Qp = (long long)((b->max_q + 32 * K - 1) / (32 * K)) * 32 * K, Rp = (long long)b->max_r + 34;
        // lowest value any stored quantity can take (gaps-only path bounds H from below; Smith-Waterman: H >= 0) and the highest score
        const long long lo = sw ? go - 2
                           : aff ? 2 * go + (Qp + Rp) * ge + open + std::min<long long>(x, 0) + ge - 2
                                 : (Qp + Rp + 1) * go + std::min<long long>(x, 0) - 2;
        const long long hi = std::max<long long>(m, 0) * std::min(Qp, Rp);
        const long long margin = 4 * std::max<long long>(std::max(-open, -ge), 1) + 16;
        B = -4 * lo + margin;
        packed = !wide && 4 * hi + B + 16 <= 32767 && !b->ctx->opt.pairwf_int32;     // else one pair per warp in int32
        if (!packed && 4 * hi + B + 16 > (1ll << 30)) return false;
        if (K == 16 && !packed) { K = 8; continue; }                               // the 16-row instantiation exists for packed Gotoh only
        break;
    }
    pl->K = K; pl->packed = packed; pl->wide = wide;
    pl->lut_lo = (uint32_t)tx; pl->lut_hi = (uint32_t)tm;
    auto pk = [&](long long v) { return packed ? (uint32_t)(v & 0xffff) * 0x00010001u : (uint32_t)v; };
    // add constants: packed halves need the always-carry compensation (high half pre-decremented), int32 takes the value itself
    auto addk = [&](long long v) { return packed ? (uint32_t)(v & 0xffff) | ((uint32_t)((v - 1) & 0xffff) << 16) : (uint32_t)v; };
    pl->ext2 = aff ? pk(4 * ge) : pk(1);
    const long long c = aff ? 4 * open : 4 * open - 2;           // (h' | 3) + c -> code 3 (Gotoh) / code 1 (linear)
    pl->addc = addk(c);
    pl->addc3 = addk(c + 3);
    pl->zero2 = pk(B + 3);
    pl->b0 = (int)(4 * open + B + code);
    pl->b1 = aff ? (int)(4 * (go + open) + B + code) : pl->b0;
    pl->bstep = sw ? 0 : (int)(4 * (aff ? ge : go));             // Smith-Waterman borders are 0 everywhere
    pl->dec_sub = (int)(sw ? B : B + code); pl->dec_add = (int)(sw ? 0 : -open);
    return true;
}

template <int ALGO, bool TB, bool PACKED, bool GBND, bool WIDE = false, int K = 8>
static int launch_pairwf_w(dpx_ctx* ctx, cudaStream_t st, const PwArgs& a, size_t smem, int n_slots) {
    auto kern = pw_nw_kernel<ALGO, TB, K, PACKED, GBND, WIDE>;
    CU(max_dyn_smem(ctx, kern));
    int per_sm = 0;
    CU(occupancy(ctx, &per_sm, kern, 128, smem));
    if (per_sm < 1) { ctx->err = "pair-wavefront kernel does not fit on an SM"; return DPX_ERR_RANGE; }
    const int blocks = std::max(1, std::min(ctx->sm_count * per_sm, (n_slots + 3) / 4));
    kern<<<blocks, 128, smem, st>>>(a);
    CU(cudaGetLastError());
    return DPX_OK;
}
template <int ALGO, bool TB>
static int launch_pairwf(dpx_ctx* ctx, cudaStream_t st, const PwArgs& a, size_t smem, int n_slots, bool packed, int K = 8) {
    if constexpr (ALGO == DPX_ALGO_ANW && !TB) {
        if (K == 16 && packed && !a.codes)
            return a.bnd_global ? launch_pairwf_w<ALGO, TB, true, true, false, 16>(ctx, st, a, smem, n_slots) : launch_pairwf_w<ALGO, TB, true, false, false, 16>(ctx, st, a, smem, n_slots);
    }
    if (a.codes) return a.bnd_global ? launch_pairwf_w<ALGO, TB, false, true, true>(ctx, st, a, smem, n_slots) : launch_pairwf_w<ALGO, TB, false, false, true>(ctx, st, a, smem, n_slots);
    if (a.bnd_global) return packed ? launch_pairwf_w<ALGO, TB, true, true>(ctx, st, a, smem, n_slots) : launch_pairwf_w<ALGO, TB, false, true>(ctx, st, a, smem, n_slots);
    return packed ? launch_pairwf_w<ALGO, TB, true, false>(ctx, st, a, smem, n_slots) : launch_pairwf_w<ALGO, TB, false, false>(ctx, st, a, smem, n_slots);
}


// ---- banded Smith-Waterman with the band mapped onto one warp (band.cuh) ---------------------------------------------
struct BandPlan { uint32_t lut_lo, lut_hi; int gadd, zerog, kb; };

static bool band_eligible(const dpx_batch* b, const dpx_params* p, int band, BandPlan* pl) {
    if (p->algo != DPX_ALGO_BSW || !b->packed2 || b->ctx->opt.no_bandkernel) return false;
    const long long m = p->match, x = p->mismatch, g = p->gap_open;
    if (!(m > 0 && x < 0 && g < 0) || band < 0 || band > 96) return false;
    const long long tm = 4 * (m - g) - 1, tx = 4 * (x - g) - 1;
    if (tm < -128 || tm > 127 || tx < -128 || tx > 127 || g < -(1 << 20)) return false;   // int8 table entries
    const long long hcmax = 4 * m * (long long)std::min(b->max_q, b->max_r) + 3;
    int nb = 0; while ((hcmax >> nb) != 0) ++nb;
    const int kb = std::min(16, 32 - nb);
    if (kb < 4) return false;
    uint8_t tab[8];
    for (int k = 0; k < 8; ++k) tab[k] = (uint8_t)tx;
    tab[3] = (uint8_t)tm;
    pl->lut_lo = tab[0] | tab[1] << 8 | tab[2] << 16 | (uint32_t)tab[3] << 24;
    pl->lut_hi = tab[4] | tab[5] << 8 | tab[6] << 16 | (uint32_t)tab[7] << 24;
    pl->gadd = (int)(4 * g - 2); pl->zerog = (int)(4 * g + 1); pl->kb = kb;
    return true;
}

template <int M, bool EXTRA>
static int launch_band(dpx_ctx* ctx, cudaStream_t st, const BandArgs& a, bool tb) {
    auto go = [&](auto kern) -> int {
        int per_sm = 0;
        CU(occupancy(ctx, &per_sm, kern, 128, 0));
        const int blocks = std::max(1, std::min(ctx->sm_count * std::max(per_sm, 1), (a.count + 3) / 4));
        kern<<<blocks, 128, 0, st>>>(a);
        CU(cudaGetLastError());
        return DPX_OK;
    };
    return tb ? go(band_sw_kernel<M, EXTRA, true>) : go(band_sw_kernel<M, EXTRA, false>);
}

static int launch_band_any(dpx_ctx* ctx, cudaStream_t st, const BandArgs& a, bool tb) {
    const BandGeom g = BandGeom::make(a.W);
    switch (g.M * 2 + g.extra) {
        case 2: return launch_band<1, false>(ctx, st, a, tb);
        case 3: return launch_band<1, true>(ctx, st, a, tb);
        case 4: return launch_band<2, false>(ctx, st, a, tb);
        case 5: return launch_band<2, true>(ctx, st, a, tb);
        case 6: return launch_band<3, false>(ctx, st, a, tb);
        default: return launch_band<3, true>(ctx, st, a, tb);
    }
}

static int batch_run(dpx_batch* b, const dpx_params* p) {
    dpx_ctx* ctx = b->ctx;
    cudaStream_t st = b->stream;
    if (p->algo < DPX_ALGO_LNW || p->algo > DPX_ALGO_ABSW) return DPX_ERR_INVALID;
    const bool banded = p->algo == DPX_ALGO_BSW || p->algo == DPX_ALGO_ABSW;
    if (banded && p->band < 0) return DPX_ERR_INVALID;
    const size_t n = b->n_pairs;
    b->params = *p; b->ran = true;
    b->stats = dpx_run_stats{};
    for (auto e : b->ev) ev_put(ctx, e, true);
    for (auto e : b->ev_sync) ev_put(ctx, e, false);
    b->ev.clear(); b->ev_kind.clear(); b->ev_sync.clear();
    const bool want_strings = (p->flags & DPX_OUT_STRINGS) != 0;
    const int algo = p->algo;
    const int CB = (algo == DPX_ALGO_ANW || algo == DPX_ALGO_ABSW) ? 4 : 2;
    const int K = (b->max_q <= 128) ? 4 : 8;
    int band = -1;
    if (banded) band = std::min(p->band, std::max(b->max_q, b->max_r));
    b->stats.cells = b->info.cells;
    b->stats.kernel_id = DPX_KERNEL_WAVEFRONT_S32;
    if (n == 0) { CU(cudaEventRecord(b->ev_begin, st)); CU(cudaEventRecord(b->ev_end, st)); return DPX_OK; }
    unsigned int* counters = ctx->counters + 64 * b->lane;   // 64 counters per lane

    // one-time (per batch) preparation, outside the timed first-kernel -> last-byte window
    { int s = ensure_order(b); if (s) return s; }
    if (want_strings) { int s = ensure_str_off(b); if (s) return s; s = ensure_blob(b); if (s) return s; }
    if (banded) {
        if (!b->d_band_cells && !pool_alloc(ctx, &b->d_band_cells, 1)) return DPX_ERR_NOMEM;
        CU(cudaMemsetAsync(b->d_band_cells, 0, sizeof(unsigned long long), st));
        band_cells_kernel<<<std::min<int>((int)((n + 7) / 8), ctx->sm_count * 8), 256, 0, st>>>(b->d_pairs, (int)n, band, b->d_band_cells);
    }
    CU(cudaEventRecord(b->ev_begin, st));

    auto add_event_pair = [&](int kind, cudaEvent_t* s, cudaEvent_t* e) -> int {
        CU(ev_get(ctx, s, true)); CU(ev_get(ctx, e, true));
        b->ev.push_back(*s); b->ev.push_back(*e); b->ev_kind.push_back(kind);
        return DPX_OK;
    };

    // ---- short-read path: packed int16x2 DPX kernel (score / end cell only) -------------------------
    {
        int B = 0, kbits = 0; bool wide = false;
        if (short_eligible(b, p, &B, &wide, &kbits)) {
            const int g = p->gap_open;
            SrArgs sa{};
            sa.packed = b->d_packed; sa.pk_off = b->d_pk_off; sa.pk_stride = b->pk_stride;
            sa.pairs = b->d_pairs; sa.order = b->d_order;
            sa.codes = wide ? b->d_codes : nullptr;
            sa.n_pairs = (int)n; sa.n_slots = wide ? (int)n : (int)((n + 1) / 2);
            const int ms = p->match - g, xs = p->mismatch - g;
            sa.ms_byte = (uint32_t)(ms & 0xff); sa.xs_byte = (uint32_t)(xs & 0xff);
            auto pk = [](int v) { return (uint32_t)(v & 0xffff) | ((uint32_t)(v & 0xffff) << 16); };
            sa.one = 1u; sa.kbits = kbits; sa.kmul = 1u << kbits; sa.B = B; sa.B2 = pk(B); sa.Bg2 = pk(B + g);
            sa.G2 = (uint32_t)(g & 0xffff) | ((uint32_t)((g - 1) & 0xffff) << 16);
            sa.scores = b->d_scores; sa.end_rc = b->d_end_rc;
            sa.counter = counters;
            CU(cudaMemsetAsync(sa.counter, 0, sizeof(unsigned int), st));
            const bool track = (p->flags & DPX_OUT_END_COORDS) != 0;
            cudaEvent_t s, e;
            { int r = add_event_pair(0, &s, &e); if (r) return r; }
            CU(cudaEventRecord(s, st));
            int r;
            if (b->max_q <= 64) r = run_short<8, 8>(ctx, b, sa, track, wide);
            else                r = run_short<8, 19>(ctx, b, sa, track, wide);
            if (r) return r;
            CU(cudaEventRecord(e, st));
            b->stats.kernel_launches = 1;
            b->stats.kernel_id = DPX_KERNEL_SHORT_S16X2;
            CU(cudaEventRecord(b->ev_end, st));
            return DPX_OK;
        }
    }

    // ---- NW / Gotoh: packed two-pair wavefront with directions in the low score bits (pairwf.cuh) ------------
    {
        PwPlan pl;
        if (pairwf_eligible(b, p, &pl)) {
            const bool aff = algo == DPX_ALGO_ANW;
            const PwGeom geo = PwGeom::make(pl.K, aff ? 4 : 2);
            const unsigned long long tbs = want_strings ? geo.words(b->max_q, b->max_r) : 0;
            // Traceback runs are cut into chunks over TWO slab buffers: the backtrack of chunk c runs on a second stream while
            // the fill kernel of chunk c+1 writes the other buffer (the walk is latency-bound and leaves the issue slots to the fill).
            size_t per_chunk = n; int nbuf = 1;
            if (want_strings) {
                const size_t ppw = pl.packed ? 2 : 1;                                    // pairs per warp
                const size_t slots_total = (n + ppw - 1) / ppw;
                const size_t max_slots = std::max<size_t>(1, (ctx->tb_budget_bytes / 2 / 4) / std::max<unsigned long long>(tbs, 1));
                size_t nchunks = (slots_total + max_slots - 1) / max_slots;
                if (!ctx->opt.serial_chunks)                                               // (set by bench.py to time the fill kernel alone)
                    nchunks = std::max<size_t>(nchunks, std::min<size_t>(8, n / 16384));        // >= 16k pairs per chunk: whole waves of warps
                const size_t slots = (slots_total + nchunks - 1) / nchunks;
                per_chunk = ppw * slots; nbuf = (nchunks > 1 && !ctx->opt.serial_chunks) ? 2 : 1;
                const size_t need = (size_t)nbuf * slots * (size_t)tbs;
                if (b->d_tb && b->tb_words < need) { CU(cudaStreamSynchronize(st)); ctx->pool.release(b->d_tb); b->d_tb = nullptr; }
                if (!b->d_tb) { if (!pool_alloc(ctx, &b->d_tb, need)) return DPX_ERR_NOMEM; b->tb_words = need; }
                b->stats.traceback_bytes = (uint64_t)slots_total * tbs * 4;
            }
            // streams of the chunk pipeline: fills alternate between the batch stream and a second one (the tail of one fill
            // overlaps the head of the next), walks run on a third
            cudaStream_t bt_st = nbuf > 1 ? ctx->aux_stream[0] : st;
            cudaStream_t fill2_st = nbuf > 1 ? ctx->aux_stream[1] : st;
            auto sync_event = [&](cudaEvent_t* ev) -> int { CU(ev_get(ctx, ev, false)); b->ev_sync.push_back(*ev); return DPX_OK; };
            std::vector<cudaEvent_t> bt_done, fill_done;
            if (nbuf > 1) {
                b->used_aux = true;
                cudaEvent_t start;
                { int r2 = sync_event(&start); if (r2) return r2; }
                CU(cudaEventRecord(start, st));
                CU(cudaStreamWaitEvent(fill2_st, start, 0));
            }
            size_t bnd_slice = 0;
            PwArgs a{};
            a.packed = b->d_packed; a.pk_off = b->d_pk_off; a.pk_stride = b->pk_stride; a.pairs = b->d_pairs; a.order = b->d_order;
            a.lut_lo = pl.lut_lo; a.lut_hi = pl.lut_hi; a.ext2 = pl.ext2; a.addc = pl.addc; a.addc3 = pl.addc3; a.minus1 = 0xffffffffu; a.zero2 = pl.zero2;
            a.one = 1u; a.four = 4u; a.sixteen = 16u;
            a.b0 = pl.b0; a.b1 = pl.b1; a.bstep = pl.bstep; a.dec_sub = pl.dec_sub; a.dec_add = pl.dec_add;
            a.scores = b->d_scores; a.end_rc = b->d_end_rc; a.tb = b->d_tb; a.tb_stride = tbs;
            a.bnd_stride = b->max_r + 36; a.rsel_stride = (b->max_r + 68) & ~1;
            a.codes = pl.wide ? b->d_codes : nullptr;
            size_t smem = (size_t)4 * a.bnd_stride * (aff ? 2 : 1) * 4 + (size_t)4 * a.rsel_stride * 2;
            if (smem > 44 * 1024) {
                // long references: the boundary rows would leave fewer than 5 blocks per SM; keep them in a per-warp global buffer
                // (one slice per slab buffer: with nbuf = 2 the fills of chunks c and c + 1 run concurrently on two streams)
                const size_t warps = (size_t)ctx->sm_count * 16 * 4;
                bnd_slice = warps * (size_t)a.bnd_stride * (aff ? 2 : 1);
                const size_t need = (size_t)nbuf * bnd_slice;
                if (ctx->boundary_ints[b->lane] < need) {
                    if (ctx->boundary[b->lane]) { CU(cudaStreamSynchronize(st)); cudaFree(ctx->boundary[b->lane]); ctx->boundary[b->lane] = nullptr; ctx->boundary_ints[b->lane] = 0; }
                    CU(cudaMalloc(&ctx->boundary[b->lane], need * sizeof(int32_t))); ctx->boundary_ints[b->lane] = need;
                }
                a.bnd_global = reinterpret_cast<uint32_t*>(ctx->boundary[b->lane]);
                smem = (size_t)4 * a.rsel_stride * 2;
            }
            b->stats.kernel_id = pl.packed ? DPX_KERNEL_PAIR_S16X2 : DPX_KERNEL_PAIR_S32;
            int c = 0;
            for (size_t first = 0; first < n; first += per_chunk, ++c) {
                a.first = (int)first; a.count = (int)std::min(per_chunk, n - first);
                a.counter = counters + (c % 64);
                a.tb = want_strings ? b->d_tb + (size_t)(c % nbuf) * (per_chunk / (pl.packed ? 2 : 1)) * (size_t)tbs : nullptr;
                if (bnd_slice) a.bnd_global = reinterpret_cast<uint32_t*>(ctx->boundary[b->lane]) + (size_t)(c % nbuf) * bnd_slice;
                cudaStream_t fst = (c & 1) ? fill2_st : st;
                if (nbuf > 1 && c >= nbuf) CU(cudaStreamWaitEvent(fst, bt_done[c - nbuf], 0));     // the buffer's previous walk is over
                CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), fst));
                cudaEvent_t s, e;
                { int r = add_event_pair(0, &s, &e); if (r) return r; }
                CU(cudaEventRecord(s, fst));
                const int n_slots = pl.packed ? (a.count + 1) / 2 : a.count;
                int r;
                if (algo == DPX_ALGO_LSW) r = launch_pairwf<DPX_ALGO_LSW, true>(ctx, fst, a, smem, n_slots, pl.packed);
                else if (aff) r = want_strings ? launch_pairwf<DPX_ALGO_ANW, true>(ctx, fst, a, smem, n_slots, pl.packed, pl.K) : launch_pairwf<DPX_ALGO_ANW, false>(ctx, fst, a, smem, n_slots, pl.packed, pl.K);
                else     r = want_strings ? launch_pairwf<DPX_ALGO_LNW, true>(ctx, fst, a, smem, n_slots, pl.packed) : launch_pairwf<DPX_ALGO_LNW, false>(ctx, fst, a, smem, n_slots, pl.packed);
                if (r) return r;
                CU(cudaEventRecord(e, fst));
                b->stats.kernel_launches++;
                if (want_strings) {
                    if (nbuf > 1) {
                        cudaEvent_t filled;
                        { int r2 = sync_event(&filled); if (r2) return r2; }
                        CU(cudaEventRecord(filled, fst));
                        CU(cudaStreamWaitEvent(bt_st, filled, 0));
                        fill_done.push_back(filled);
                    }
                    PwBtArgs t{};
                    t.blob = b->d_blob; t.pairs = b->d_pairs; t.order = a.order; t.first = a.first; t.count = a.count; t.K = pl.K;
                    t.tb = a.tb; t.tb_stride = tbs; t.strings = b->d_strings; t.str_off = b->d_str_off; t.str_start = b->d_str_start;
                    t.scores = b->d_scores; t.end_rc = b->d_end_rc;
                    { int r2 = add_event_pair(1, &s, &e); if (r2) return r2; }
                    CU(cudaEventRecord(s, bt_st));
                    // one thread walks one pair: small launches use one-warp blocks so that the walkers spread over every SM
                    const int bt_threads = a.count < ctx->sm_count * 512 ? 32 : 128;
                    const int bt_blocks = (a.count + bt_threads - 1) / bt_threads;
                    if (pl.packed) {
                        if (algo == DPX_ALGO_LSW) pw_bt_kernel<DPX_ALGO_LSW, 8, true><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                        else if (aff) pw_bt_kernel<DPX_ALGO_ANW, 8, true><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                        else     pw_bt_kernel<DPX_ALGO_LNW, 8, true><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                    } else {
                        if (algo == DPX_ALGO_LSW) pw_bt_kernel<DPX_ALGO_LSW, 8, false><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                        else if (aff) pw_bt_kernel<DPX_ALGO_ANW, 8, false><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                        else     pw_bt_kernel<DPX_ALGO_LNW, 8, false><<<bt_blocks, bt_threads, 0, bt_st>>>(t);
                    }
                    CU(cudaGetLastError());
                    CU(cudaEventRecord(e, bt_st));
                    b->stats.kernel_launches++;
                    if (nbuf > 1) {
                        cudaEvent_t walked;
                        { int r2 = sync_event(&walked); if (r2) return r2; }
                        CU(cudaEventRecord(walked, bt_st));
                        bt_done.push_back(walked);
                    }
                }
            }
            for (size_t k = 0; k < bt_done.size(); ++k) CU(cudaStreamWaitEvent(st, bt_done[k], 0));     // the batch stream ends after every walk
            CU(cudaEventRecord(b->ev_end, st));
            return DPX_OK;
        }
    }

    // ---- banded SW: the band mapped onto one warp (band.cuh) ----------------------------------------------------
    {
        BandPlan pl;
        if (band_eligible(b, p, band, &pl)) {
            const BandGeom geo = BandGeom::make(band);
            const int qs_len = geo.qs_len(b->max_q, b->max_r), rs_len = geo.rs_len(b->max_q, b->max_r);
            if (b->band_prep_w != band) {
                if (b->d_band_qs) { CU(cudaStreamSynchronize(st)); ctx->pool.release(b->d_band_qs); ctx->pool.release(b->d_band_rs); b->d_band_qs = b->d_band_rs = nullptr; }
                if (!pool_alloc(ctx, &b->d_band_qs, n * (size_t)qs_len) || !pool_alloc(ctx, &b->d_band_rs, n * (size_t)rs_len)) return DPX_ERR_NOMEM;
                band_prep_kernel<<<std::min<int>((int)((n + 7) / 8), ctx->sm_count * 8), 256, 0, st>>>(
                    b->d_packed, b->d_pk_off, b->pk_stride, b->d_pairs, (int)n, geo.offq(), geo.offr(), qs_len, rs_len, b->d_band_qs, b->d_band_rs);
                CU(cudaGetLastError());
                b->band_prep_w = band;
                CU(cudaEventRecord(b->ev_begin, st));       // stream layout is per-batch preparation, like the 2-bit pack
            }
            const unsigned long long tbs = want_strings ? geo.words(b->max_q, b->max_r) : 0;
            size_t per_chunk = n;
            if (want_strings) {
                per_chunk = std::max<size_t>(1, std::min<size_t>(n, (ctx->tb_budget_bytes / 4) / std::max<unsigned long long>(tbs, 1)));
                if (b->d_tb && b->tb_words < per_chunk * (size_t)tbs) { CU(cudaStreamSynchronize(st)); ctx->pool.release(b->d_tb); b->d_tb = nullptr; }
                if (!b->d_tb) { if (!pool_alloc(ctx, &b->d_tb, per_chunk * (size_t)tbs)) return DPX_ERR_NOMEM; b->tb_words = per_chunk * (size_t)tbs; }
                b->stats.traceback_bytes = (uint64_t)n * tbs * 4;
            }
            BandArgs a{};
            a.pairs = b->d_pairs; a.order = b->d_order; a.qs = b->d_band_qs; a.rs = b->d_band_rs; a.qs_len = qs_len; a.rs_len = rs_len;
            a.W = band; a.lut_lo = pl.lut_lo; a.lut_hi = pl.lut_hi; a.gadd = pl.gadd; a.zerog = pl.zerog; a.kb = pl.kb; a.kmul = 1u << pl.kb;
            a.one = 1u; a.four = 4u; a.sixteen = 16u; a.minus1 = 0xffffffffu; a.scores = b->d_scores; a.end_rc = b->d_end_rc; a.tb = b->d_tb; a.tb_stride = tbs;
            b->stats.kernel_id = DPX_KERNEL_BAND_S32;
            int c = 0;
            for (size_t first = 0; first < n; first += per_chunk, ++c) {
                a.first = (int)first; a.count = (int)std::min(per_chunk, n - first);
                a.counter = counters + (c % 64);
                CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
                cudaEvent_t s, e;
                { int r = add_event_pair(0, &s, &e); if (r) return r; }
                CU(cudaEventRecord(s, st));
                { int r = launch_band_any(ctx, st, a, want_strings); if (r) return r; }
                CU(cudaEventRecord(e, st));
                b->stats.kernel_launches++;
                if (want_strings) {
                    BandBtArgs t{};
                    t.blob = b->d_blob; t.blob_lo = b->d_blob_alloc; t.pairs = b->d_pairs; t.order = a.order; t.first = a.first; t.count = a.count; t.W = band;
                    t.scores = b->d_scores; t.end_rc = b->d_end_rc; t.tb = b->d_tb; t.tb_stride = tbs;
                    t.strings = b->d_strings; t.str_off = b->d_str_off; t.str_start = b->d_str_start;
                    { int r2 = add_event_pair(1, &s, &e); if (r2) return r2; }
                    CU(cudaEventRecord(s, st));
                    t.wshift = band_bt_wpw_shift(a.count, ctx->sm_count);
                    const int wpw = 1 << t.wshift;
                    const size_t bt_smem = (size_t)band_bt_smem_words(geo.M, wpw) * sizeof(uint32_t);
                    CU(max_dyn_smem(ctx, band_bt_kernel));
                    band_bt_kernel<<<(a.count + wpw - 1) / wpw, 32, bt_smem, st>>>(t);
                    CU(cudaGetLastError());
                    CU(cudaEventRecord(e, st));
                    b->stats.kernel_launches++;
                }
            }
            CU(cudaEventRecord(b->ev_end, st));
            return DPX_OK;
        }
    }

    // ---- general path: warp-per-pair wavefront (+ traceback and GPU backtrack) ----------------------------
    { int s = ensure_blob(b); if (s) return s; }              // byte-compare kernels: sidecar batches materialise their bytes first
    const unsigned long long tb_stride = want_strings ? WfGeom::make(K, CB, b->max_q, b->max_r, band).words() : 0;
    size_t per_chunk = n;
    if (want_strings) {
        per_chunk = std::max<size_t>(1, std::min<size_t>(n, (ctx->tb_budget_bytes / 4) / std::max<unsigned long long>(tb_stride, 1)));
        if (b->d_tb && b->tb_words < per_chunk * (size_t)tb_stride) { CU(cudaStreamSynchronize(st)); ctx->pool.release(b->d_tb); b->d_tb = nullptr; }
        if (!b->d_tb) { if (!pool_alloc(ctx, &b->d_tb, per_chunk * (size_t)tb_stride)) return DPX_ERR_NOMEM; b->tb_words = per_chunk * (size_t)tb_stride; }
        b->stats.traceback_bytes = (uint64_t)n * tb_stride * 4;
    }

    WfArgs a{};
    a.blob = b->d_blob; a.pairs = b->d_pairs; a.order = b->d_order;
    a.match = p->match; a.mismatch = p->mismatch; a.go = p->gap_open; a.ge = p->gap_extend; a.band = band;
    a.scores = b->d_scores; a.end_rc = b->d_end_rc;
    a.tb = want_strings ? b->d_tb : nullptr; a.tb_stride = tb_stride;
    a.rmax_p1 = b->max_r + 1;
    a.boundary_stride = 2LL * (b->max_r + 1);
    int blocks = 0;
    { int r = dispatch_wf(ctx, st, algo, want_strings, K, a, 0, true, (int)std::min<size_t>(n, 1u << 30), &blocks); if (r) return r; }
    const size_t need = (size_t)blocks * 4 * (size_t)a.boundary_stride;
    if (ctx->boundary_ints[b->lane] < need) {
        if (ctx->boundary[b->lane]) { CU(cudaStreamSynchronize(st)); cudaFree(ctx->boundary[b->lane]); ctx->boundary[b->lane] = nullptr; ctx->boundary_ints[b->lane] = 0; }
        CU(cudaMalloc(&ctx->boundary[b->lane], need * sizeof(int32_t))); ctx->boundary_ints[b->lane] = need;
    }
    a.boundary = ctx->boundary[b->lane];

    int c = 0;
    for (size_t first = 0; first < n; first += per_chunk, ++c) {
        a.first = (int)first; a.count = (int)std::min(per_chunk, n - first);
        a.counter = counters + (c % 64);
        CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
        cudaEvent_t s, e;
        { int r = add_event_pair(0, &s, &e); if (r) return r; }
        CU(cudaEventRecord(s, st));
        dispatch_wf(ctx, st, algo, want_strings, K, a, blocks, false, 0, nullptr);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e, st));
        b->stats.kernel_launches++;
        if (want_strings) {
            BtArgs t{};
            t.blob = b->d_blob; t.pairs = b->d_pairs; t.order = a.order; t.first = a.first; t.count = a.count;
            t.K = K; t.band = band; t.scores = b->d_scores; t.end_rc = b->d_end_rc; t.tb = b->d_tb; t.tb_stride = tb_stride;
            t.strings = b->d_strings; t.str_off = b->d_str_off; t.str_start = b->d_str_start;
            { int r = add_event_pair(1, &s, &e); if (r) return r; }
            CU(cudaEventRecord(s, st));
            const int bt_blocks = (a.count + 127) / 128;
            switch (algo) {
                case DPX_ALGO_LNW: bt_walk_kernel<DPX_ALGO_LNW><<<bt_blocks, 128, 0, st>>>(t); break;
                case DPX_ALGO_ANW: bt_walk_kernel<DPX_ALGO_ANW><<<bt_blocks, 128, 0, st>>>(t); break;
                case DPX_ALGO_LSW: bt_walk_kernel<DPX_ALGO_LSW><<<bt_blocks, 128, 0, st>>>(t); break;
                case DPX_ALGO_BSW: bt_walk_kernel<DPX_ALGO_BSW><<<bt_blocks, 128, 0, st>>>(t); break;
                case DPX_ALGO_ABSW: bt_walk_kernel<DPX_ALGO_ABSW><<<bt_blocks, 128, 0, st>>>(t); break;
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(e, st));
            b->stats.kernel_launches++;
        }
    }
    CU(cudaEventRecord(b->ev_end, st));
    return DPX_OK;
}

// D2H of scores / end cells into caller memory at their final place; asynchronous.
static int batch_fetch_async(dpx_batch* b, int32_t* scores, int32_t* end_rc) {
    dpx_ctx* ctx = b->ctx;
    const size_t n = b->n_pairs;
    if (scores && n) CU(cudaMemcpyAsync(scores, b->d_scores, n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    if (end_rc && n) CU(cudaMemcpyAsync(end_rc, b->d_end_rc, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    return DPX_OK;
}

