// host_allmax.cuh — host side of the all-maxima LinearSmithWaterman mode (csrc/allmax.cuh); included inside dpxalign.cu's extern "C" block.
#pragma once

int dpx_align_batch_text_all(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                             const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                             char** text, size_t* text_bytes, long long* n_alignments) {
    if (!ctx || !params || !text || !text_bytes || (!sequences && n_bytes) || (!pairs && n_pairs) || n_pairs > 0x7fffffffu) return DPX_ERR_INVALID;
    if (params->algo != DPX_ALGO_LSW) return DPX_ERR_UNSUPPORTED;
    *text = nullptr; *text_bytes = 0; if (n_alignments) *n_alignments = 0;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = n_pairs;
    // whole score matrices: 4 B per cell, all pairs back to back
    std::vector<long long> moff(n + 1, 0);
    for (size_t k = 0; k < n; ++k) {
        const dpx_seq_pair& pr = pairs[k];
        if (pr.referenceSize < 0 || pr.querySize < 0 || pr.referenceIdx < 0 || pr.queryIdx < 0 ||
            (size_t)pr.referenceIdx + (size_t)pr.referenceSize > n_bytes || (size_t)pr.queryIdx + (size_t)pr.querySize > n_bytes) return DPX_ERR_INVALID;
        moff[k + 1] = moff[k] + ((long long)pr.referenceSize + 1) * ((long long)pr.querySize + 1);
    }
    const long long cells = moff[n];
    if (cells > (3LL << 30)) { ctx->err = "all-maxima mode keeps whole score matrices: more than 12 GiB for this batch"; return DPX_ERR_RANGE; }
    uint8_t *d_blob = nullptr, *d_text = nullptr; dpx_seq_pair* d_pairs = nullptr; int32_t *d_H = nullptr, *d_best = nullptr, *d_count = nullptr;
    long long *d_moff = nullptr, *d_soff = nullptr, *d_moves = nullptr, *d_toff = nullptr; int *d_si = nullptr, *d_sj = nullptr, *d_sp = nullptr;
    char* host = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        DevPool& P = ctx->pool;
        P.release(d_blob); P.release(d_text); P.release(d_pairs); P.release(d_H); P.release(d_best); P.release(d_count); P.release(d_moff);
        P.release(d_soff); P.release(d_moves); P.release(d_toff); P.release(d_si); P.release(d_sj); P.release(d_sp);
    };
#define ACU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); if (host) dpx_free(host); return DPX_ERR_CUDA; } } while (0)
    std::vector<int32_t> best(n, 0), count(n, 0);
    std::vector<long long> soff(n + 1, 0);
    long long S = 0;
    AmArgs a{};
    if (n) {
        if (!pool_alloc(ctx, &d_blob, n_bytes + 16) || !pool_alloc(ctx, &d_pairs, n) || !pool_alloc(ctx, &d_H, (size_t)cells) || !pool_alloc(ctx, &d_best, n) ||
            !pool_alloc(ctx, &d_count, n) || !pool_alloc(ctx, &d_moff, n + 1) || !pool_alloc(ctx, &d_soff, n + 1)) { cleanup(); return DPX_ERR_NOMEM; }
        ACU(cudaMemcpyAsync(d_blob, sequences, n_bytes, cudaMemcpyHostToDevice, st));
        ACU(cudaMemcpyAsync(d_pairs, pairs, sizeof(dpx_seq_pair) * n, cudaMemcpyHostToDevice, st));
        ACU(cudaMemcpyAsync(d_moff, moff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
        ACU(cudaMemsetAsync(d_H, 0, sizeof(int32_t) * (size_t)cells, st));                 // row 0 and column 0 of every matrix
        a.blob = d_blob; a.pairs = d_pairs; a.n_pairs = (int)n; a.match = params->match; a.mismatch = params->mismatch; a.gap = params->gap_open;
        a.H = d_H; a.moff = d_moff; a.best = d_best; a.count = d_count; a.soff = d_soff;
        const int warp_blocks = (int)((n + 3) / 4);
        am_fill_kernel<<<warp_blocks, 128, 0, st>>>(a);
        ACU(cudaGetLastError());
        ACU(cudaMemcpyAsync(best.data(), d_best, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        ACU(cudaMemcpyAsync(count.data(), d_count, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        ACU(cudaStreamSynchronize(st));
        for (size_t k = 0; k < n; ++k) soff[k + 1] = soff[k] + count[k];
        S = soff[n];
    }
    std::vector<long long> moves((size_t)S, 0), toff((size_t)S, 0);
    if (S) {
        if (!pool_alloc(ctx, &d_si, (size_t)S) || !pool_alloc(ctx, &d_sj, (size_t)S) || !pool_alloc(ctx, &d_sp, (size_t)S) ||
            !pool_alloc(ctx, &d_moves, (size_t)S) || !pool_alloc(ctx, &d_toff, (size_t)S)) { cleanup(); return DPX_ERR_NOMEM; }
        ACU(cudaMemcpyAsync(d_soff, soff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
        a.start_i = d_si; a.start_j = d_sj; a.start_pair = d_sp; a.n_starts = S; a.moves = d_moves; a.toff = d_toff;
        am_list_kernel<<<(int)((n + 3) / 4), 128, 0, st>>>(a);
        am_walk_kernel<<<(int)((S + 127) / 128), 128, 0, st>>>(a);
        ACU(cudaGetLastError());
        ACU(cudaMemcpyAsync(moves.data(), d_moves, sizeof(long long) * (size_t)S, cudaMemcpyDeviceToHost, st));
        ACU(cudaStreamSynchronize(st));
    }
    // text layout: "<i> | <score>\n", then per alignment REF / REL / QRY (moves + 1 bytes each), alignments of a pair ordered by
    // (moves, queue order); score 0: three empty lines (c++/LinearSmithWaterman.cpp:253-257)
    std::vector<size_t> head_off(n, 0);
    std::vector<std::string> heads(n);
    size_t total = 0;
    std::vector<long long> order;
    for (size_t k = 0; k < n; ++k) {
        heads[k] = std::to_string(first_index + (long long)k) + " | " + std::to_string(best[k]) + "\n";
        if (best[k] == 0) heads[k] += "\n\n\n";
        head_off[k] = total; total += heads[k].size();
        order.resize((size_t)count[k]);
        for (long long m = 0; m < count[k]; ++m) order[(size_t)m] = soff[k] + m;
        std::stable_sort(order.begin(), order.end(), [&](long long x, long long y) { return moves[(size_t)x] < moves[(size_t)y]; });
        for (long long s : order) { toff[(size_t)s] = (long long)total; total += (size_t)(3 * (moves[(size_t)s] + 1)); }
    }
    host = (char*)g_host.take(std::max<size_t>(total + 1, 1));
    if (!host) { cleanup(); return DPX_ERR_NOMEM; }
    if (S) {
        if (!pool_alloc(ctx, &d_text, total + 16)) { cleanup(); dpx_free(host); return DPX_ERR_NOMEM; }
        ACU(cudaMemcpyAsync(d_toff, toff.data(), sizeof(long long) * (size_t)S, cudaMemcpyHostToDevice, st));
        a.text = d_text;
        am_emit_kernel<<<(int)((S + 127) / 128), 128, 0, st>>>(a);
        ACU(cudaGetLastError());
        ACU(cudaMemcpyAsync(host, d_text, total, cudaMemcpyDeviceToHost, st));
        ACU(cudaStreamSynchronize(st));
    }
#undef ACU
    for (size_t k = 0; k < n; ++k) memcpy(host + head_off[k], heads[k].data(), heads[k].size());
    host[total] = 0;
    cleanup();
    *text = host; *text_bytes = total; if (n_alignments) *n_alignments = S;
    return DPX_OK;
}
