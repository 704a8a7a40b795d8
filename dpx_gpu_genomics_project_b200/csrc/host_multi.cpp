// host_multi.cpp — multi-GPU mode A behind the C ABI (dpx_create_multi / dpx_multi_align_batch, include/dpxalign.h):
// one host process drives every GPU of the node.  Plain host C++ written ONLY against the public ABI (each worker owns a
// dpx_ctx and calls dpx_align_batch on its shard), so everything the single-GPU call does — sidecar upload, chunk pipeline,
// kernel selection — is what runs on every device.
//
// The reference has no multi-GPU code (no cudaSetDevice anywhere); its CPU driver splits the pair list over
// pthreads in consecutive blocks (c++/main.cpp:166-232), which is the shape kept here: CONTIGUOUS shards, balanced by
// cell count.  Contiguous (rather than dealt) shards let every GPU upload one slice of the blob / packed sidecar instead of
// a gather; length bucketing happens inside each shard on the device (sched_keys_kernel + radix sort).
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dpxalign.h"

namespace {

struct Worker {
    int device = 0;
    dpx_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, stop = false, done = true;
    int create_status = DPX_OK;

    void loop() {
        dpx_bind_host_to_device(device);          // page-locked staging allocated by this thread lands next to its GPU
        create_status = dpx_create(&ctx, device);
        { std::lock_guard<std::mutex> g(mu); done = true; }
        cv.notify_all();
        for (;;) {
            std::function<void()> j;
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return has_job || stop; });
                if (stop && !has_job) break;
                j = std::move(job); has_job = false;
            }
            j();
            { std::lock_guard<std::mutex> g(mu); done = true; }
            cv.notify_all();
        }
        if (ctx) { dpx_destroy(ctx); ctx = nullptr; }
    }
    void post(std::function<void()> j) {
        { std::lock_guard<std::mutex> g(mu); job = std::move(j); has_job = true; done = false; }
        cv.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return done && !has_job; }); }
};

}  // namespace

struct dpx_multi {
    std::vector<Worker*> w;
    std::string err;
    std::mutex call_mu;                            // one batch call at a time per dpx_multi
};

extern "C" {

int dpx_multi_shard_bounds(const dpx_seq_pair* pairs, size_t n_pairs, int n_shards, size_t* bounds) {
    if ((!pairs && n_pairs) || n_shards < 1 || !bounds) return DPX_ERR_INVALID;
    // weight of a pair = its cell count (at least 1, so that empty pairs are spread as well)
    auto weight = [&](size_t i) -> unsigned long long {
        const long long r = pairs[i].referenceSize, q = pairs[i].querySize;
        const unsigned long long c = (r > 0 && q > 0) ? (unsigned long long)r * (unsigned long long)q : 0ull;
        return c ? c : 1ull;
    };
    // prefix sums over a few host threads (millions of pairs: a single-threaded pass would cost as much as the alignment itself):
    // per-range totals first, then every boundary is located inside the one range its target falls in
    const int T = n_pairs >= (1u << 18) ? (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1;
    std::vector<unsigned long long> part((size_t)T + 1, 0);
    auto range = [&](int t) { return n_pairs * (size_t)t / (size_t)T; };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) {
            auto fn = [&, t] { unsigned long long a = 0; for (size_t i = range(t); i < range(t + 1); ++i) a += weight(i); part[(size_t)t + 1] = a; };
            if (T == 1) fn(); else th.emplace_back(fn);
        }
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) part[(size_t)t + 1] += part[(size_t)t];
    const unsigned long long total = part[(size_t)T];
    bounds[0] = 0;
    for (int g = 1; g < n_shards; ++g) {
        // first index i such that the weight before i, plus half of pair i, reaches the target (same rule as a sequential scan)
        const unsigned long long target = (unsigned long long)((long double)total * g / n_shards);
        int t = 0;
        while (t + 1 < T && part[(size_t)t + 1] < target) ++t;
        unsigned long long acc = part[(size_t)t]; size_t i = range(t);
        while (i < n_pairs && acc + weight(i) / 2 < target) acc += weight(i++);
        bounds[g] = std::max(i, bounds[g - 1]);
    }
    bounds[n_shards] = n_pairs;
    return DPX_OK;
}

// Boundaries for a call: a registered input whose pairs all have the same lengths (the parser knows) splits evenly without
// touching the index; anything else takes the prefix-sum pass.
static void shard_bounds_for(const char* sequences, const dpx_seq_pair* pairs, size_t n_pairs, int n_shards, size_t* bounds) {
    size_t reg_pairs = 0;
    if (sequences && dpx_input_sidecar(sequences, &reg_pairs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) == 2) {
        for (int g = 0; g <= n_shards; ++g) bounds[g] = n_pairs * (size_t)g / (size_t)n_shards;
        return;
    }
    dpx_multi_shard_bounds(pairs, n_pairs, n_shards, bounds);
}

void dpx_destroy_multi(dpx_multi* m) {
    if (!m) return;
    for (Worker* x : m->w) {
        { std::lock_guard<std::mutex> g(x->mu); x->stop = true; }
        x->cv.notify_all();
        if (x->th.joinable()) x->th.join();
        delete x;
    }
    delete m;
}

int dpx_create_multi(dpx_multi** out, const int* devices, int n_devices) {
    if (!out || n_devices < 1) return DPX_ERR_INVALID;
    *out = nullptr;
    const int have = dpx_device_count();
    if (have <= 0) return DPX_ERR_NO_DEVICE;
    for (int g = 0; g < n_devices; ++g) {
        const int d = devices ? devices[g] : g;
        if (d < 0 || d >= have) return DPX_ERR_INVALID;
    }
    dpx_multi* m = new dpx_multi();
    for (int g = 0; g < n_devices; ++g) {
        Worker* x = new Worker();
        x->device = devices ? devices[g] : g;
        x->done = false;
        m->w.push_back(x);
        x->th = std::thread([x] { x->loop(); });
    }
    int st = DPX_OK;
    for (Worker* x : m->w) { x->wait(); if (x->create_status != DPX_OK && st == DPX_OK) st = x->create_status; }
    if (st != DPX_OK) { dpx_destroy_multi(m); return st; }
    // Chunks per worker and call (dpx_multi_set_option overrides).  A process that owns 4 GPUs alone runs the 12-chunk pipeline of
    // dpx_align_batch best (tools/multi_trace.py 4 4000000: 3.52 ms per 4 M pairs at 12 chunks, 3.91 at 3, 5.21 at 1).  At 8 devices the
    // only measurement is from inside a torchrun job (bench.py, `2_multi_abi`): 17.3 ms at 12 chunks, 4.29 at 3 -- ~400 driver calls
    // per worker and call from 8 threads of one process do not stay cheap -- so from 8 devices up a worker takes few, large chunks.
    if (n_devices >= 8)
        for (Worker* x : m->w) { dpx_set_option(x->ctx, "chunks_packed", 3); dpx_set_option(x->ctx, "chunks", 4); }
    *out = m;
    return DPX_OK;
}

int dpx_multi_device_count(const dpx_multi* m) { return m ? (int)m->w.size() : 0; }
const char* dpx_multi_last_error(const dpx_multi* m) { return m ? m->err.c_str() : "null dpx_multi"; }

int dpx_multi_set_option(dpx_multi* m, const char* name, long long value) {
    if (!m) return DPX_ERR_INVALID;
    std::lock_guard<std::mutex> g(m->call_mu);
    for (Worker* x : m->w) { const int st = dpx_set_option(x->ctx, name, value); if (st) { m->err = dpx_last_error(x->ctx); return st; } }
    return DPX_OK;
}

int dpx_multi_align_batch(dpx_multi* m, const dpx_params* params, const char* sequences, size_t n_bytes,
                          const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                          char** strings_blob, size_t** string_offsets) {
    if (!m || !params || !scores || (!sequences && n_bytes) || (!pairs && n_pairs)) return DPX_ERR_INVALID;
    std::lock_guard<std::mutex> lock(m->call_mu);
    const bool want_strings = (params->flags & DPX_OUT_STRINGS) && strings_blob && string_offsets;
    if (strings_blob) *strings_blob = nullptr;
    if (string_offsets) *string_offsets = nullptr;
    const int G = (int)m->w.size();
    std::vector<size_t> bounds((size_t)G + 1);
    shard_bounds_for(sequences, pairs, n_pairs, G, bounds.data());
    std::vector<int> status((size_t)G, DPX_OK);
    std::vector<char*> sb((size_t)G, nullptr); std::vector<size_t*> so((size_t)G, nullptr);
    std::vector<size_t> sbytes((size_t)G, 0);
    // phase 1: every device aligns its shard; scores / end cells land at their final place
    for (int g = 0; g < G; ++g) {
        Worker* x = m->w[(size_t)g];
        const size_t b0 = bounds[(size_t)g], n = bounds[(size_t)g + 1] - b0;
        x->post([=, &status, &sb, &so, &sbytes] {
            if (n == 0) return;
            status[(size_t)g] = dpx_align_batch(x->ctx, params, sequences, n_bytes, pairs + b0, n, scores + b0,
                                                end_row_col ? end_row_col + 2 * b0 : nullptr,
                                                want_strings ? &sb[(size_t)g] : nullptr, want_strings ? &so[(size_t)g] : nullptr);
            if (status[(size_t)g] == DPX_OK && want_strings) {
                const size_t last = so[(size_t)g][3 * n - 1];
                sbytes[(size_t)g] = last + strlen(sb[(size_t)g] + last) + 1;
            }
        });
    }
    int st = DPX_OK;
    for (int g = 0; g < G; ++g) {
        m->w[(size_t)g]->wait();
        if (status[(size_t)g] != DPX_OK && st == DPX_OK) { st = status[(size_t)g]; m->err = "device " + std::to_string(m->w[(size_t)g]->device) + ": " + dpx_last_error(m->w[(size_t)g]->ctx); }
    }
    auto drop = [&] { for (int g = 0; g < G; ++g) { dpx_free(sb[(size_t)g]); dpx_free(so[(size_t)g]); } };
    if (st != DPX_OK || !want_strings) { drop(); return st; }
    // phase 2: stitch the per-shard string blobs into one blob + one offset table, each worker copying its own part
    std::vector<size_t> base((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) base[(size_t)g + 1] = base[(size_t)g] + sbytes[(size_t)g];
    char* blob = (char*)malloc(base[(size_t)G] ? base[(size_t)G] : 1);
    size_t* offs = (size_t*)malloc((n_pairs ? 3 * n_pairs : 1) * sizeof(size_t));
    if (!blob || !offs) { free(blob); free(offs); drop(); return DPX_ERR_NOMEM; }
    for (int g = 0; g < G; ++g) {
        Worker* x = m->w[(size_t)g];
        const size_t b0 = bounds[(size_t)g], n = bounds[(size_t)g + 1] - b0, at = base[(size_t)g];
        x->post([=, &sb, &so, &sbytes] {
            if (n == 0) return;
            memcpy(blob + at, sb[(size_t)g], sbytes[(size_t)g]);
            for (size_t k = 0; k < 3 * n; ++k) offs[3 * b0 + k] = so[(size_t)g][k] + at;
        });
    }
    for (int g = 0; g < G; ++g) m->w[(size_t)g]->wait();
    drop();
    *strings_blob = blob; *string_offsets = offs;
    return DPX_OK;
}

int dpx_multi_align_batch_text(dpx_multi* m, const dpx_params* params, const char* sequences, size_t n_bytes,
                               const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                               int32_t* scores, int32_t* end_row_col, char** text, size_t* text_bytes) {
    if (!m || !params || !text || !text_bytes || (!sequences && n_bytes) || (!pairs && n_pairs)) return DPX_ERR_INVALID;
    std::lock_guard<std::mutex> lock(m->call_mu);
    *text = nullptr; *text_bytes = 0;
    const int G = (int)m->w.size();
    std::vector<size_t> bounds((size_t)G + 1);
    shard_bounds_for(sequences, pairs, n_pairs, G, bounds.data());
    std::vector<int> status((size_t)G, DPX_OK);
    std::vector<char*> tx((size_t)G, nullptr); std::vector<size_t> tb((size_t)G, 0);
    for (int g = 0; g < G; ++g) {
        Worker* x = m->w[(size_t)g];
        const size_t b0 = bounds[(size_t)g], n = bounds[(size_t)g + 1] - b0;
        x->post([=, &status, &tx, &tb] {
            if (n == 0) return;
            status[(size_t)g] = dpx_align_batch_text(x->ctx, params, sequences, n_bytes, pairs + b0, n, first_index + (long long)b0,
                                                     scores ? scores + b0 : nullptr, end_row_col ? end_row_col + 2 * b0 : nullptr,
                                                     &tx[(size_t)g], &tb[(size_t)g]);
        });
    }
    int st = DPX_OK;
    for (int g = 0; g < G; ++g) {
        m->w[(size_t)g]->wait();
        if (status[(size_t)g] != DPX_OK && st == DPX_OK) { st = status[(size_t)g]; m->err = "device " + std::to_string(m->w[(size_t)g]->device) + ": " + dpx_last_error(m->w[(size_t)g]->ctx); }
    }
    auto drop = [&] { for (int g = 0; g < G; ++g) dpx_free(tx[(size_t)g]); };
    if (st != DPX_OK) { drop(); return st; }
    std::vector<size_t> base((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) base[(size_t)g + 1] = base[(size_t)g] + tb[(size_t)g];
    char* out = (char*)malloc(base[(size_t)G] + 1);
    if (!out) { drop(); return DPX_ERR_NOMEM; }
    for (int g = 0; g < G; ++g) {
        Worker* x = m->w[(size_t)g];
        const size_t at = base[(size_t)g];
        x->post([=, &tx, &tb] { if (tb[(size_t)g]) memcpy(out + at, tx[(size_t)g], tb[(size_t)g]); });
    }
    for (int g = 0; g < G; ++g) m->w[(size_t)g]->wait();
    drop();
    out[base[(size_t)G]] = 0;
    *text = out; *text_bytes = base[(size_t)G];
    return DPX_OK;
}

}  // extern "C"
