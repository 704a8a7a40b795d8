// host_long_abi.cuh — C ABI of the long-pair path (dpx_align_long_pair, dpx_stripe_*), included inside dpxalign.cu's
// extern "C" block.
#pragma once

int dpx_stripe_create(dpx_ctx* ctx, const dpx_params* params, const char* ref_stripe, size_t R_local, size_t col_offset,
                      const char* qry, size_t Q, int stripe_index, int n_stripes, dpx_stripe** out) {
    if (!ctx || !params || !out || !ref_stripe || !qry || R_local == 0 || Q == 0 || n_stripes < 1 || stripe_index < 0 || stripe_index >= n_stripes) return DPX_ERR_INVALID;
    if (params->algo != DPX_ALGO_LSW) return DPX_ERR_UNSUPPORTED;
    if (Q > 0x7ffffff0u || (long double)params->match * (long double)Q > 2.0e9L) return DPX_ERR_RANGE;
    CU(cudaSetDevice(ctx->device));
    dpx_stripe* s = new dpx_stripe();
    s->ctx = ctx; s->params = *params; s->R_local = R_local; s->col_offset = col_offset; s->Q = Q; s->index = stripe_index; s->n = n_stripes;
    auto fail = [&](int st) { dpx_stripe_free(s); return st; };
    // Every stripe must realise the reference's relation, plain byte equality (c++/LinearSmithWaterman.cpp:92), whatever
    // kernel it picks, so that per-rank choices cannot disagree.  The table kernel codes bytes with a FIXED injective map
    // of ONE family -- digits '0'..'3', or upper-case ACGT, or lower-case acgt -- and is only taken when the whole query
    // and this stripe's reference lie inside that one family (code equality == byte equality there).  Anything else
    // (a fifth symbol, mixed case such as a soft-masked reference against upper-case reads, digits next to letters) sends
    // this stripe to the byte-compare kernel, which is the same relation by construction.
    static const auto fixed_code = [](uint8_t c) -> int {
        switch (c) { case '0': case 'A': case 'a': return 0; case '1': case 'C': case 'c': return 1;
                     case '2': case 'G': case 'g': return 2; case '3': case 'T': case 't': return 3; default: return -1; }
    };
    static const auto family = [](uint8_t c) -> int { return c <= '9' ? 1 : (c < 'a' ? 2 : 4); };
    bool table = long_table_ok(params, Q, Q) && !ctx->opt.long_notable;
    int fam = 0;
    for (size_t i = 0; i < Q && table; ++i) { const uint8_t c = (uint8_t)qry[i]; if (fixed_code(c) < 0) table = false; fam |= family(c); }
    for (size_t i = 0; i < R_local && table; ++i) { const uint8_t c = (uint8_t)ref_stripe[i]; if (fixed_code(c) < 0) table = false; fam |= family(c); }
    if (fam & (fam - 1)) table = false;           // two families: 'a' and 'A' (or '0' and 'A') would share a code
    s->mode = ((long double)params->match * (long double)Q < 8.0e6L && params->match > 0 ? 1 : 0) | (table ? 2 : 0);
    // lane width: the whole stripe must be one co-resident pass
    // The right edge a stripe exports is the last column of its last warp, so every stripe but the last must be made of WHOLE
    // warps: its width has to be a multiple of 32 K.
    const bool last = stripe_index == n_stripes - 1;
    auto whole = [&](int k) { return last || R_local % (32ull * (unsigned)k) == 0; };
    const bool allow32 = table && (long double)params->match * (long double)Q < 6.0e7L && whole(32);
    int K = long_pick_k(ctx, (long long)R_local, allow32), cap = 0;
    if (const int k = ctx->opt.long_k) { if (k != 32 || allow32) K = k; }
    while (K > 2 && !whole(K)) K /= 2;
    if (!whole(K)) { ctx->err = "a stripe that is not the last one must be a multiple of 64 columns wide"; return fail(DPX_ERR_INVALID); }
    for (;;) {
        if (long_capacity_k(ctx, K, s->mode, &cap)) return fail(DPX_ERR_CUDA);
        if ((long long)((R_local + 32ull * K - 1) / (32ull * K)) <= cap || K == (allow32 ? 32 : 16) || !whole(2 * K)) break;
        K *= 2;
    }
    s->K = K; s->nw = (int)((R_local + 32ull * K - 1) / (32ull * K));
    if (s->nw > cap) { ctx->err = "stripe too wide for one co-resident pass"; return fail(DPX_ERR_RANGE); }
    const size_t nw = (size_t)s->nw;
    bool ok = cudaMalloc(&s->xbuf, 128 + sizeof(unsigned long long) * (size_t)dpx_stripe::XRING) == cudaSuccess &&
              cudaMalloc(&s->d_ref, R_local + 16) == cudaSuccess && cudaMalloc(&s->d_qry, Q + 16) == cudaSuccess &&
              cudaMalloc(&s->d_rings, sizeof(unsigned long long) * nw * (size_t)dpx_stripe::RING) == cudaSuccess &&
              cudaMalloc(&s->d_cnt, sizeof(long long) * 2 * (nw + 2)) == cudaSuccess &&
              cudaMalloc(&s->d_bs, sizeof(int32_t) * nw) == cudaSuccess && cudaMalloc(&s->d_br, sizeof(long long) * nw) == cudaSuccess &&
              cudaMalloc(&s->d_bc, sizeof(long long) * nw) == cudaSuccess && cudaMalloc(&s->d_chans, sizeof(LongChan) * (nw + 1)) == cudaSuccess &&
              cudaMalloc(&s->d_err, sizeof(int)) == cudaSuccess &&
              cudaEventCreate(&s->e0) == cudaSuccess && cudaEventCreate(&s->e1) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return fail(DPX_ERR_NOMEM); }
    std::vector<uint8_t> cref, cqry;
    if (table) {
        cref.resize(R_local); cqry.resize(Q);
        for (size_t i = 0; i < R_local; ++i) cref[i] = (uint8_t)fixed_code((uint8_t)ref_stripe[i]);
        for (size_t i = 0; i < Q; ++i) cqry[i] = (uint8_t)fixed_code((uint8_t)qry[i]);
    }
    if (cudaMemcpy(s->d_ref, table ? (const char*)cref.data() : ref_stripe, R_local, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->d_qry, table ? (const char*)cqry.data() : qry, Q, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(s->xbuf, 0, 128) != cudaSuccess) return fail(DPX_ERR_CUDA);
    *out = s;
    return DPX_OK;
}

int dpx_stripe_export(dpx_stripe* s, void* handle) {
    if (!s || !handle) return DPX_ERR_INVALID;
    dpx_ctx* ctx = s->ctx;
    static_assert(sizeof(cudaIpcMemHandle_t) == DPX_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->xbuf));
    memcpy(handle, &h, sizeof(h));
    return DPX_OK;
}

int dpx_stripe_connect(dpx_stripe* s, const void* prev_handle, const void* next_handle) {
    if (!s) return DPX_ERR_INVALID;
    dpx_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    if (prev_handle && s->index > 0) {
        cudaIpcMemHandle_t h; memcpy(&h, prev_handle, sizeof(h));
        CU(cudaIpcOpenMemHandle((void**)&s->prev_x, h, cudaIpcMemLazyEnablePeerAccess));
    }
    if (next_handle && s->index + 1 < s->n) {
        cudaIpcMemHandle_t h; memcpy(&h, next_handle, sizeof(h));
        CU(cudaIpcOpenMemHandle((void**)&s->next_x, h, cudaIpcMemLazyEnablePeerAccess));
    }
    if ((s->index > 0 && !s->prev_x) || (s->index + 1 < s->n && !s->next_x)) return DPX_ERR_INVALID;
    // channels: [0] = inbox (own memory; credit goes back to prev), [1..nw-1] local rings, [nw] = next stripe's inbox (peer)
    const long long nw = s->nw;
    std::vector<LongChan> ch((size_t)nw + 1);
    long long* cred = s->d_cnt;
    for (long long c = 0; c <= nw; ++c) {
        LongChan x{};
        if (c == 0) {
            if (s->index > 0) { x.ring = (unsigned long long*)(s->xbuf + 128); x.size = dpx_stripe::XRING; x.credit = (long long*)(s->prev_x + 8); }
        } else if (c == nw) {
            if (s->index + 1 < s->n) { x.ring = (unsigned long long*)(s->next_x + 128); x.size = dpx_stripe::XRING; x.credit = (long long*)(s->xbuf + 8); }
        } else { x.ring = s->d_rings + (c - 1) * dpx_stripe::RING; x.size = dpx_stripe::RING; x.credit = cred + c; }
        ch[(size_t)c] = x;
    }
    CU(cudaMemcpy(s->d_chans, ch.data(), sizeof(LongChan) * (size_t)(nw + 1), cudaMemcpyHostToDevice));
    return DPX_OK;
}

int dpx_stripe_reset(dpx_stripe* s) {
    if (!s) return DPX_ERR_INVALID;
    dpx_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemset(s->xbuf, 0, 128 + sizeof(unsigned long long) * (size_t)dpx_stripe::XRING));     // credit + inbox tags
    CU(cudaMemset(s->d_cnt, 0, sizeof(long long) * 2 * ((size_t)s->nw + 2)));
    CU(cudaMemset(s->d_rings, 0, sizeof(unsigned long long) * (size_t)s->nw * (size_t)dpx_stripe::RING));
    CU(cudaMemset(s->d_err, 0, sizeof(int)));
    CU(cudaDeviceSynchronize());
    return DPX_OK;
}

int dpx_stripe_run(dpx_stripe* s) {
    if (!s) return DPX_ERR_INVALID;
    dpx_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    LongArgs a{};
    a.ref = s->d_ref; a.qry = s->d_qry; a.Q = (long long)s->Q; a.R_local = (long long)s->R_local; a.col0 = 0; a.col_offset = (long long)s->col_offset;
    a.match = s->params.match; a.mismatch = s->params.mismatch; a.gap = s->params.gap_open; a.nwarps = s->nw; a.chans = s->d_chans;
    a.best_score = s->d_bs; a.best_row = s->d_br; a.best_col = s->d_bc; a.error_flag = s->d_err; a.system_scope = s->n > 1;
    a.tab_match = s->params.match - s->params.gap_open; a.tab_mismatch = s->params.mismatch - s->params.gap_open; a.sixteen = 16u;
    CU(cudaEventRecord(s->e0, ctx->stream));
    { int st = long_launch_k(ctx, s->K, s->mode, a, ctx->stream); if (st) return st; }
    CU(cudaEventRecord(s->e1, ctx->stream));
    s->launched = true;
    return DPX_OK;
}

int dpx_stripe_result(dpx_stripe* s, int32_t* score, int64_t* end_row, int64_t* end_col, double* kernel_ms) {
    if (!s || !s->launched || !score) return DPX_ERR_INVALID;
    dpx_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t nw = (size_t)s->nw;
    std::vector<int32_t> bs(nw); std::vector<long long> br(nw), bc(nw); int err = 0;
    CU(cudaMemcpy(bs.data(), s->d_bs, sizeof(int32_t) * nw, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(br.data(), s->d_br, sizeof(long long) * nw, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(bc.data(), s->d_bc, sizeof(long long) * nw, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { ctx->err = "stripe pipeline watchdog fired (a neighbour never produced / consumed)"; return DPX_ERR_CUDA; }
    int32_t best = 0; long long r0 = 0, c0 = 0;
    for (size_t w = 0; w < nw; ++w)
        if (bs[w] > best || (bs[w] == best && best > 0 && (br[w] < r0 || (br[w] == r0 && bc[w] < c0)))) { best = bs[w]; r0 = br[w]; c0 = bc[w]; }
    *score = best; if (end_row) *end_row = r0; if (end_col) *end_col = c0;
    if (kernel_ms) { float ms = 0; CU(cudaEventElapsedTime(&ms, s->e0, s->e1)); *kernel_ms = ms; }
    return DPX_OK;
}

void dpx_stripe_free(dpx_stripe* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaDeviceSynchronize();
    if (s->prev_x) cudaIpcCloseMemHandle(s->prev_x);
    if (s->next_x) cudaIpcCloseMemHandle(s->next_x);
    cudaFree(s->xbuf); cudaFree(s->d_ref); cudaFree(s->d_qry); cudaFree(s->d_rings); cudaFree(s->d_cnt); cudaFree(s->d_bs);
    cudaFree(s->d_br); cudaFree(s->d_bc); cudaFree(s->d_chans); cudaFree(s->d_err);
    if (s->e0) cudaEventDestroy(s->e0);
    if (s->e1) cudaEventDestroy(s->e1);
    delete s;
}

int dpx_align_long_pair(dpx_ctx* ctx, const dpx_params* params, const char* ref, size_t R, const char* qry, size_t Q,
                        int32_t* score, int64_t* end_row, int64_t* end_col) {
    if (!ctx || !params || !score || (!ref && R) || (!qry && Q)) return DPX_ERR_INVALID;
    if (params->algo != DPX_ALGO_LSW) return DPX_ERR_UNSUPPORTED;
    CU(cudaSetDevice(ctx->device));
    *score = 0; if (end_row) *end_row = 0; if (end_col) *end_col = 0;
    if (R == 0 || Q == 0) return DPX_OK;
    if ((long double)params->match * (long double)std::min(R, Q) > 2.0e9L || Q > 0x7ffffff0u || R > 0x7ffffff0u) return DPX_ERR_RANGE;     // int32 scores / rows
    return long_pair_single(ctx, params, ref, R, qry, Q, score, end_row, end_col);
}

int dpx_align_long_pair_strings(dpx_ctx* ctx, const dpx_params* params, const char* ref, size_t R, const char* qry, size_t Q,
                                int32_t* score, int64_t* end_row, int64_t* end_col, int64_t* start_row, int64_t* start_col,
                                char** lines, size_t* line_len, double* stage_ms) {
    if (!ctx || !params || !score || !lines || !line_len || (!ref && R) || (!qry && Q)) return DPX_ERR_INVALID;
    if (params->algo != DPX_ALGO_LSW) return DPX_ERR_UNSUPPORTED;
    CU(cudaSetDevice(ctx->device));
    *score = 0; *lines = nullptr; *line_len = 0;
    if (end_row) *end_row = 0; if (end_col) *end_col = 0; if (start_row) *start_row = 0; if (start_col) *start_col = 0;
    if (stage_ms) for (int k = 0; k < 6; ++k) stage_ms[k] = 0;
    if (R == 0 || Q == 0) {
        char* blob = (char*)g_host.take(4);
        if (!blob) return DPX_ERR_NOMEM;
        blob[0] = blob[1] = blob[2] = 0; *lines = blob;
        return DPX_OK;
    }
    if ((long double)params->match * (long double)std::min(R, Q) > 2.0e9L || Q > 0x7ffffff0u || R > 0x7ffffff0u) return DPX_ERR_RANGE;
    LongTraceStats ls;
    const int r = long_pair_strings(ctx, params, ref, R, qry, Q, score, end_row, end_col, start_row, start_col, lines, line_len, &ls);
    if (r == DPX_OK && stage_ms) { stage_ms[0] = ls.fwd_ms; stage_ms[1] = (double)ls.rounds; stage_ms[2] = ls.walk_ms; stage_ms[3] = (double)ls.tiles; stage_ms[4] = ls.TH; stage_ms[5] = ls.TW; }
    return r;
}
