// dpxalign.cu — C ABI (include/dpxalign.h) and host-side orchestration of libdpxalign.so.
// sm_100a only; no CPU fallback: every alignment entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <map>
#include <unordered_map>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"
#include "host_pack.h"
#include "dpx_ops.cuh"
#include "wavefront.cuh"
#include "pack.cuh"
#include "backtrack.cuh"
#include "shortread.cuh"
#include "longpair.cuh"
#include "longtrace.cuh"
#include "allmax.cuh"
#include "pairwf.cuh"
#include "band.cuh"

using namespace dpx;

// ------------------------------------------------------------------------------------------------
// Grow-only caching allocator for device memory: dpx_align_batch is called once per batch in the drop-in
// driver, so cudaMalloc/cudaFree must stay off its critical path.
struct DevPool {
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> size_;
    void* alloc(size_t bytes) {
        bytes = std::max<size_t>((bytes + 511) & ~(size_t)511, 512);
        auto it = free_.lower_bound(bytes);
        if (it != free_.end() && it->first <= bytes * 4 + (1u << 20)) { void* p = it->second; free_.erase(it); return p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            clear_free();                                   // give cached blocks back and retry once
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        size_[p] = bytes;
        return p;
    }
    void release(void* p) { if (!p) return; auto it = size_.find(p); if (it != size_.end()) free_.emplace(it->second, p); }
    void clear_free() { for (auto& kv : free_) { cudaFree(kv.second); size_.erase(kv.second); } free_.clear(); }
    void clear_all() { for (auto& kv : size_) cudaFree(kv.first); size_.clear(); free_.clear(); }
};

// ------------------------------------------------------------------------------------------------
// Host blobs the library hands out (strings_blob / string_offsets) are page-locked, so their D2H copy runs at PCIe
// speed instead of through the driver's pageable staging; dpx_free returns them to this process-wide cache (bounded),
// so a driver that aligns batch after batch does not pay cudaHostAlloc each time.  Anything not allocated here (the
// parser's malloc'ed arrays) still goes to free().
#include <atomic>
#include <mutex>
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> live_;
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCached = (size_t)2 << 30;     // documented in dpxalign.h (dpx_free / dpx_trim)
    void* take(size_t bytes) {
        bytes = std::max<size_t>((bytes + 4095) & ~(size_t)4095, 4096);
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_.lower_bound(bytes);
            if (it != free_.end() && it->first <= 2 * bytes + (1u << 20)) {
                void* p = it->second; const size_t sz = it->first; free_.erase(it); cached_bytes -= sz; live_[p] = sz; return p;
            }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        std::lock_guard<std::mutex> g(mu);
        live_[p] = bytes;
        return p;
    }
    bool give_back(void* p) {
        std::unique_lock<std::mutex> g(mu);
        auto it = live_.find(p);
        if (it == live_.end()) return false;
        const size_t sz = it->second; live_.erase(it);
        if (cached_bytes + sz <= kMaxCached) { free_.emplace(sz, p); cached_bytes += sz; return true; }
        g.unlock();
        cudaFreeHost(p);
        return true;
    }
    // gives every cached (not handed-out) block back to the driver
    void trim() {
        std::multimap<size_t, void*> drop;
        { std::lock_guard<std::mutex> g(mu); drop.swap(free_); cached_bytes = 0; }
        for (auto& kv : drop) cudaFreeHost(kv.second);
    }
};
static HostCache g_host;
static std::atomic<int> g_live_ctx{0};

// Debug / test knobs, set only through dpx_set_option (never read from the environment).
struct DpxOptions {
    int long_k = 0;            // long pair: force the lane width (2, 4, 8, 16, 32); 0 = cost model
    int long_cap = 0;          // long pair: cap on co-resident warps (forces several passes); 0 = occupancy
    int long_notable = 0;      // long pair: byte-compare kernel even for <= 4 symbols
    int long_bt_tiles = 0;     // long-pair traceback: tiles per round; 0 = 2 per SM
    int no_pairwf = 0;         // skip the packed pair-wavefront kernels (general wavefront instead)
    int pairwf_int32 = 0;      // pair-wavefront kernels in int32 even when int16x2 would fit
    int pairwf_k8 = 0;         // Gotoh pair-wavefront kernel with 8 rows per lane even where 16 would be taken
    int no_bandkernel = 0;     // skip the band-on-a-warp kernel
    int no_shortread = 0;      // skip the short-read kernel
    int serial_chunks = 0;     // traceback chunk pipeline: one buffer, chunks in series (times the fill kernel alone)
    int trace = 0;             // one-call pipeline: per-chunk timeline on stderr
    int no_sidecar = 0;        // dpx_align_batch: ignore the parser's packed sidecar (upload the raw blob)
    int serial_strings = 0;    // dpx_align_batch with strings: one batch, upload -> kernels -> compaction -> download in series
};

struct dpx_ctx {
    int device = 0;
    DpxOptions opt;
    int sm_count = 0;
    // opt-in ceiling of dynamic shared memory per block.  Kernels always get THIS as their MaxDynamicSharedMemorySize attribute:
    // the attribute is per function and device, not per launch, so two contexts on one device (one per host thread) setting
    // each launch's own size would race (thread A's larger launch after thread B's smaller attribute = invalid argument).
    int smem_optin = 48 * 1024;
    cudaStream_t own_stream = nullptr;             // lane 0 (unless the caller supplies a stream)
    cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};   // lanes 1..3 of the chunked one-call pipeline
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;            // every H2D input copy goes through this one stream: strict chunk order on the DMA engine
    std::string err;
    DevPool pool;
    BatchInfo* h_info[4] = {nullptr, nullptr, nullptr, nullptr};     // pinned read-back slots, one per lane
    int32_t* boundary[4] = {nullptr, nullptr, nullptr, nullptr}; size_t boundary_ints[4] = {0, 0, 0, 0};
    unsigned int* counters = nullptr;              // 64 dynamic-work counters per lane
    size_t tb_budget_bytes = (size_t)16 << 30;     // traceback chunk budget
    // recycled CUDA events (creating / destroying ~10 events per chunk of the one-call pipeline was a measurable share of its
    // host time, and with several contexts in one process those calls serialise on the driver's lock)
    std::vector<cudaEvent_t> ev_pool_timing, ev_pool_plain;
    // per kernel: shared-memory attribute already raised, and occupancy per dynamic shared-memory size
    std::unordered_map<const void*, std::map<size_t, int>> occ_cache;
    int chunks = 12;                               // one-call pipeline: equal middle chunks (dpx_set_option "chunks")
    int chunks_strings = 8;                        // ... upper limit of chunks when alignment strings are requested
    int chunks_packed = 12;                        // ... when the input comes from the packed sidecar (kernel-bound: fewer, larger chunks)
    bool counted = false;                          // registered in g_live_ctx (dpx_create succeeded)
};

struct dpx_batch {
    dpx_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    int lane = 0;
    size_t n_pairs = 0;
    long long byte_lo = 0, byte_hi = 0;
    BatchInfo info{};
    int max_r = 0, max_q = 0, min_r = 0, min_q = 0;
    bool uniform = true;
    int n_symbols = 0;
    bool packed2 = false;
    PackLut lut;
    // device inputs
    uint8_t* d_blob_alloc = nullptr;               // holds bytes [byte_lo, byte_hi)
    const uint8_t* d_blob = nullptr;               // d_blob_alloc - byte_lo: indexable with the seqPair offsets
    dpx_seq_pair* d_pairs = nullptr;
    uint32_t* d_packed = nullptr;
    uint32_t inv4 = 0;                             // sidecar batches: code -> byte (4 bytes), for ensure_blob
    bool from_sidecar = false;
    uint32_t* d_stage[2] = {nullptr, nullptr};     // sidecar batches: staged sizes / word offsets (released with the batch)
    std::vector<void*> d_scratch;                  // scratch of the batch's CUB passes (sort keys, scan temporaries): released with the batch
    size_t h2d_bytes = 0;                          // bytes this batch's upload moved over PCIe
    uint8_t* d_codes = nullptr;                    // 5..8 symbols: byte codes, same indexing as d_blob (allocated at d_codes_alloc - byte_lo)
    uint8_t* d_codes_alloc = nullptr;
    unsigned long long* d_pk_off = nullptr; unsigned long long pk_stride = 0;
    unsigned long long* d_str_len = nullptr;       // 3*(Q+R+1) per pair (scanned lazily into d_str_off)
    int32_t* d_order = nullptr;
    // device outputs
    int32_t* d_scores = nullptr;
    int32_t* d_end_rc = nullptr;
    uint32_t* d_tb = nullptr; size_t tb_words = 0;   // traceback slab (one or two chunk buffers) and its capacity
    std::vector<cudaEvent_t> ev_sync;                // fill-done / backtrack-done events of the chunk pipeline
    char* d_strings = nullptr;
    unsigned long long* d_str_off = nullptr;
    int32_t* d_str_start = nullptr;
    unsigned long long* d_band_cells = nullptr;
    uint8_t* d_band_qs = nullptr; uint8_t* d_band_rs = nullptr; int band_prep_w = -1;   // padded streams of the band kernel (band.cuh)
    BatchInfo* d_info = nullptr;
    // run state
    bool ran = false; dpx_params params{};
    bool used_aux = false;           // a run queued work on the ctx's auxiliary streams (dpx_batch_free drains them too)
    std::vector<cudaEvent_t> ev;     // pairs of (start, end) per kernel; kind in ev_kind
    std::vector<int> ev_kind;        // 0 fill, 1 backtrack
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_h2d = nullptr;
    dpx_run_stats stats{};
};

// Kernels always get the opt-in ceiling (minus their static shared memory) as MaxDynamicSharedMemorySize: see dpx_ctx::smem_optin.
// Done once per kernel and context (the attribute never changes afterwards).
template <typename F>
static cudaError_t max_dyn_smem(dpx_ctx* ctx, F kern) {
    if (ctx->occ_cache.count((const void*)kern)) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) ctx->occ_cache[(const void*)kern];
    return e;
}
// cudaOccupancyMaxActiveBlocksPerMultiprocessor, remembered per (kernel, dynamic shared memory)
template <typename F>
static cudaError_t occupancy(dpx_ctx* ctx, int* per_sm, F kern, int threads, size_t smem) {
    auto& m = ctx->occ_cache[(const void*)kern];
    auto it = m.find(smem);
    if (it != m.end()) { *per_sm = it->second; return cudaSuccess; }
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, threads, smem);
    if (e == cudaSuccess) m[smem] = *per_sm;
    return e;
}
static cudaError_t ev_get(dpx_ctx* ctx, cudaEvent_t* ev, bool timing) {
    auto& pool = timing ? ctx->ev_pool_timing : ctx->ev_pool_plain;
    if (!pool.empty()) { *ev = pool.back(); pool.pop_back(); return cudaSuccess; }
    return timing ? cudaEventCreate(ev) : cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
}
static void ev_put(dpx_ctx* ctx, cudaEvent_t ev, bool timing) {
    if (!ev) return;
    auto& pool = timing ? ctx->ev_pool_timing : ctx->ev_pool_plain;
    if (pool.size() < 256) pool.push_back(ev); else cudaEventDestroy(ev);
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return e__ == cudaErrorMemoryAllocation ? DPX_ERR_NOMEM : DPX_ERR_CUDA;                \
        }                                                                                          \
    } while (0)

extern "C" {

int dpx_abi_version(void) { return DPX_ABI_VERSION; }

const char* dpx_strerror(int s) {
    switch (s) {
        case DPX_OK: return "ok";
        case DPX_ERR_INVALID: return "invalid argument";
        case DPX_ERR_NO_DEVICE: return "no CUDA device (libdpxalign has no CPU fallback)";
        case DPX_ERR_CUDA: return "CUDA error";
        case DPX_ERR_NOMEM: return "out of memory";
        case DPX_ERR_IO: return "cannot open or read input file";
        case DPX_ERR_FORMAT: return "number of lines not a multiple of 3";
        case DPX_ERR_RANGE: return "value out of the kernel's range";
        case DPX_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

int dpx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void dpx_destroy(dpx_ctx* ctx);

int dpx_create(dpx_ctx** out, int device) {
    if (!out) return DPX_ERR_INVALID;
    *out = nullptr;
    int n = dpx_device_count();
    if (n <= 0) return DPX_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return DPX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return DPX_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DPX_ERR_CUDA;
    dpx_ctx* ctx = new dpx_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&ctx->counters, 4 * 64 * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int l = 0; l < 3 && ok; ++l) ok = cudaStreamCreateWithFlags(&ctx->aux_stream[l], cudaStreamNonBlocking) == cudaSuccess;
    for (int l = 0; l < 4 && ok; ++l) ok = cudaHostAlloc(&ctx->h_info[l], sizeof(BatchInfo), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) { cudaGetLastError(); dpx_destroy(ctx); return DPX_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->tb_budget_bytes = std::min<size_t>((size_t)48 << 30, fr / 3);
    *out = ctx;
    ctx->counted = true;
    g_live_ctx.fetch_add(1);
    return DPX_OK;
}

void dpx_trim(void) { g_host.trim(); }

void dpx_destroy(dpx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->pool.clear_all();
    for (auto e : ctx->ev_pool_timing) cudaEventDestroy(e);
    for (auto e : ctx->ev_pool_plain) cudaEventDestroy(e);
    for (int l = 0; l < 4; ++l) { if (ctx->boundary[l]) cudaFree(ctx->boundary[l]); if (ctx->h_info[l]) cudaFreeHost(ctx->h_info[l]); }
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int l = 0; l < 3; ++l) if (ctx->aux_stream[l]) cudaStreamDestroy(ctx->aux_stream[l]);
    if (ctx->counted && g_live_ctx.fetch_sub(1) == 1) g_host.trim();       // last context gone: release the cached page-locked blocks
    delete ctx;
}

const char* dpx_last_error(const dpx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int dpx_set_stream(dpx_ctx* ctx, void* s) {
    if (!ctx) return DPX_ERR_INVALID;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return DPX_OK;
}

int dpx_set_option(dpx_ctx* ctx, const char* name, long long value) {
    if (!ctx || !name) return DPX_ERR_INVALID;
    const std::string k(name);
    DpxOptions& o = ctx->opt;
    if (k == "long_k") { if (value != 0 && value != 2 && value != 4 && value != 8 && value != 16 && value != 32) return DPX_ERR_INVALID; o.long_k = (int)value; }
    else if (k == "long_cap") { if (value < 0 || value > (1 << 20)) return DPX_ERR_INVALID; o.long_cap = (int)value; }
    else if (k == "long_notable") o.long_notable = value != 0;
    else if (k == "long_bt_tiles") { if (value < 0 || value > 4096) return DPX_ERR_INVALID; o.long_bt_tiles = (int)value; }
    else if (k == "no_pairwf") o.no_pairwf = value != 0;
    else if (k == "pairwf_int32") o.pairwf_int32 = value != 0;
    else if (k == "pairwf_k8") o.pairwf_k8 = value != 0;
    else if (k == "no_bandkernel") o.no_bandkernel = value != 0;
    else if (k == "no_shortread") o.no_shortread = value != 0;
    else if (k == "serial_chunks") o.serial_chunks = value != 0;
    else if (k == "trace") o.trace = value != 0;
    else if (k == "no_sidecar") o.no_sidecar = value != 0;
    else if (k == "serial_strings") o.serial_strings = value != 0;
    else if (k == "chunks") { if (value < 1 || value > 64) return DPX_ERR_INVALID; ctx->chunks = (int)value; }
    else if (k == "chunks_strings") { if (value < 1 || value > 64) return DPX_ERR_INVALID; ctx->chunks_strings = (int)value; }
    else if (k == "chunks_packed") { if (value < 1 || value > 64) return DPX_ERR_INVALID; ctx->chunks_packed = (int)value; }
    else if (k == "tb_budget_bytes") { if (value < (1 << 16)) return DPX_ERR_INVALID; ctx->tb_budget_bytes = (size_t)value; }
    else { ctx->err = "unknown option: " + k; return DPX_ERR_INVALID; }
    return DPX_OK;
}

void dpx_free(void* p) {
    if (!p) return;
    dpxhost_pack::forget(p);                       // a parser blob / index: its packed sidecar goes with it
    if (!g_host.give_back(p)) free(p);
}

// ---- parser (replaces c++/parseInput.cpp:9-119) -------------------------------------------------
// buf holds the file image (got bytes, newlines in place) and becomes the blob; on success it is owned by the caller.
static int parse_buffer(char* buf, size_t got, dpx_seq_pair** pairs_out, dpx_input_info* info) {
    if (got > 0x7fffffffull) return DPX_ERR_RANGE;                   // seqPair offsets are int (parseInput.h:22-29)
    size_t lines = 0;
    for (size_t i = 0; i < got; ++i) lines += (buf[i] == '\n');
    if (lines % 3 != 0) return DPX_ERR_FORMAT;                       // parseInput.cpp:38-41
    size_t n = lines / 3;
    const size_t cap = 10000000;                                     // INPUT_CAP, parseInput.cpp:7,102-105
    dpx_seq_pair* idx = (dpx_seq_pair*)malloc(std::max<size_t>(n, 1) * sizeof(dpx_seq_pair));
    if (!idx) return DPX_ERR_NOMEM;
    dpx_input_info in{}; in.minReferenceLength = SIZE_MAX; in.minQueryLength = SIZE_MAX;
    int mode = 0; size_t k = 0;
    for (size_t i = 0; i < got && k < n; ++i) {
        if (buf[i] != '\n') continue;
        buf[i] = '\0';
        if (mode == 0) { idx[k].referenceIdx = (int32_t)(i + 1); mode = 1; }
        else if (mode == 1) {
            idx[k].referenceSize = (int32_t)(i - (size_t)idx[k].referenceIdx);
            in.avgReferenceLength += idx[k].referenceSize;
            in.maxReferenceLength = std::max(in.maxReferenceLength, (size_t)idx[k].referenceSize);
            in.minReferenceLength = std::min(in.minReferenceLength, (size_t)idx[k].referenceSize);
            idx[k].queryIdx = (int32_t)(i + 1); mode = 2;
        } else {
            idx[k].querySize = (int32_t)(i - (size_t)idx[k].queryIdx);
            in.avgQueryLength += idx[k].querySize;
            in.maxQueryLength = std::max(in.maxQueryLength, (size_t)idx[k].querySize);
            in.minQueryLength = std::min(in.minQueryLength, (size_t)idx[k].querySize);
            in.numCells += (size_t)idx[k].referenceSize * (size_t)idx[k].querySize;
            ++k; mode = 0;
            if (k == cap) break;
        }
    }
    in.numPairs = k; in.numBytes = got;
    if (k) { in.avgReferenceLength /= (double)k; in.avgQueryLength /= (double)k; }
    *pairs_out = idx;
    if (info) *info = in;
    // 2-bit sidecar for the one-call path (host_pack.h); inputs with a fifth symbol simply stay unregistered
    if (k) dpxhost_pack::register_input(buf, got, idx, k);
    return DPX_OK;
}

int dpx_parse_input(const char* path, dpx_seq_pair** pairs_out, char** seq_out, dpx_input_info* info) {
    if (!path || !pairs_out || !seq_out) return DPX_ERR_INVALID;
    *pairs_out = nullptr; *seq_out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return DPX_ERR_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return DPX_ERR_IO; }
    long sz = ftell(f);
    if (sz < 0) { fclose(f); return DPX_ERR_IO; }
    rewind(f);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return DPX_ERR_NOMEM; }
    size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz) { free(buf); return DPX_ERR_IO; }
    const int st = parse_buffer(buf, got, pairs_out, info);
    if (st) { free(buf); return st; }
    *seq_out = buf;
    return DPX_OK;
}

int dpx_parse_image(const char* image, size_t n_bytes, dpx_seq_pair** pairs_out, char** seq_out, dpx_input_info* info) {
    if ((!image && n_bytes) || !pairs_out || !seq_out) return DPX_ERR_INVALID;
    *pairs_out = nullptr; *seq_out = nullptr;
    char* buf = (char*)malloc(n_bytes + 1);
    if (!buf) return DPX_ERR_NOMEM;
    if (n_bytes) memcpy(buf, image, n_bytes);
    const int st = parse_buffer(buf, n_bytes, pairs_out, info);
    if (st) { free(buf); return st; }
    *seq_out = buf;
    return DPX_OK;
}

int dpx_register_input(const char* sequences, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs) {
    if ((!sequences && n_bytes) || (!pairs && n_pairs)) return DPX_ERR_INVALID;
    return dpxhost_pack::register_input(sequences, n_bytes, pairs, n_pairs);
}

int dpx_input_sidecar(const char* sequences, size_t* n_pairs, size_t* n_words, const uint32_t** words, const uint32_t** word_offsets,
                      int* n_symbols, int* page_locked, unsigned char* code_to_byte) {
    const dpxhost_pack::Sidecar* s = sequences ? dpxhost_pack::find_blob(sequences) : nullptr;
    if (!s) return 0;
    if (n_pairs) *n_pairs = s->n_pairs;
    if (n_words) *n_words = s->n_words;
    if (words) *words = s->words;
    if (word_offsets) *word_offsets = s->woff;
    if (n_symbols) *n_symbols = s->nsym;
    if (page_locked) *page_locked = s->pinned ? 1 : 0;
    if (code_to_byte) memcpy(code_to_byte, s->inv, 4);
    return s->uniform ? 2 : 1;
}

void dpx_unregister_input(const char* sequences) { if (sequences) dpxhost_pack::forget(sequences); }

int dpx_bind_host_to_device(int device) { return dpxhost_pack::bind_thread_to_device(device); }

#include "host_fastx.cuh"

// ---- DPX instruction evaluation -----------------------------------------------------------------
int dpx_dpx_eval(dpx_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, int n,
                 uint32_t* out, uint8_t* pred_hi, uint8_t* pred_lo) {
    if (!ctx || !a || !b || !c || !out || !pred_hi || !pred_lo || n < 0 || op < 0 || op >= OP_COUNT) return DPX_ERR_INVALID;
    if (n == 0) return DPX_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t *da = nullptr, *db = nullptr, *dc = nullptr, *dout = nullptr; uint8_t *dh = nullptr, *dl = nullptr;
    auto run = [&]() -> int {
        CU(cudaMalloc(&da, 4 * (size_t)n)); CU(cudaMalloc(&db, 4 * (size_t)n)); CU(cudaMalloc(&dc, 4 * (size_t)n));
        CU(cudaMalloc(&dout, 4 * (size_t)n)); CU(cudaMalloc(&dh, n)); CU(cudaMalloc(&dl, n));
        CU(cudaMemcpyAsync(da, a, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(db, b, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dc, c, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        dpx_eval_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(op, da, db, dc, n, dout, dh, dl);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(pred_hi, dh, n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(pred_lo, dl, n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return DPX_OK;
    };
    const int st = run();
    if (st) cudaStreamSynchronize(ctx->stream);
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dout); cudaFree(dh); cudaFree(dl);
    return st;
}

int dpx_selftest_dpx(dpx_ctx* ctx) {
    // A few of the reference's known answers (c++/testFakeDPX.cpp:10-113); tests/ replays the full set.
    struct V { int op; uint32_t a, b, c, want; int ph, pl; };
    static const V vec[] = {
        {OP_VIMAX3_S32, 1, 2, 3, 3, 0, 0}, {OP_VIMAX3_S32, (uint32_t)-5, (uint32_t)-10, (uint32_t)-30, (uint32_t)-5, 0, 0},
        {OP_VIMAX3_S16X2, 0xFFFD00FF, 0xFFFE00FF, 0xFFFFFF00, 0xFFFF00FF, 0, 0},
        {OP_VIMAX3_S32_RELU, (uint32_t)-5, (uint32_t)-10, (uint32_t)-30, 0, 0, 0},
        {OP_VIMAX3_S16X2_RELU, 0, 0xFFFF00FF, 0xFFFFFF00, 0x000000FF, 0, 0},
        {OP_VIBMAX_S32, (uint32_t)-10, (uint32_t)-30, 0, (uint32_t)-10, 1, 0}, {OP_VIBMAX_S32, 1, 2, 0, 2, 0, 0},
        {OP_VIBMAX_S16X2, 0xFFFF00FF, 0xFFFFFF00, 0, 0xFFFF00FF, 1, 1}, {OP_VIBMAX_S16X2, 0xFFFD00FF, 0xFFFE01FF, 0, 0xFFFE01FF, 0, 0},
        {OP_VIADDMAX_S32, 2, 3, 1, 5, 0, 0}, {OP_VIADDMAX_S32, (uint32_t)-5, (uint32_t)-10, (uint32_t)-30, (uint32_t)-15, 0, 0},
    };
    const int n = (int)(sizeof(vec) / sizeof(vec[0]));
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t o; uint8_t h, l;
        int st = dpx_dpx_eval(ctx, vec[i].op, &vec[i].a, &vec[i].b, &vec[i].c, 1, &o, &h, &l);
        if (st != DPX_OK) return st;
        const bool two = vec[i].op >= OP_VIBMAX_S16X2 && vec[i].op <= OP_VIBMIN_U16X2;
        const bool one = vec[i].op >= OP_VIBMAX_S32 && vec[i].op <= OP_VIBMIN_U32;
        if (o != vec[i].want || ((one || two) && h != vec[i].ph) || (two && l != vec[i].pl)) ++bad;
    }
    return bad;
}

// ---- batch ----------------------------------------------------------------------------------------
}  // extern "C"

#define CUB_(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return fail(e__ == cudaErrorMemoryAllocation ? DPX_ERR_NOMEM : DPX_ERR_CUDA); } } while (0)

static void batch_release(dpx_batch* b) {
    // returns device memory to the pool; the batch's stream must have drained (callers guarantee it)
    DevPool& P = b->ctx->pool;
    P.release(b->d_blob_alloc); P.release(b->d_pairs); P.release(b->d_packed); P.release(b->d_codes_alloc); P.release(b->d_pk_off); P.release(b->d_str_len);
    P.release(b->d_order); P.release(b->d_scores); P.release(b->d_end_rc); P.release(b->d_tb); P.release(b->d_strings);
    P.release(b->d_stage[0]); P.release(b->d_stage[1]);
    for (void* x : b->d_scratch) P.release(x);
    P.release(b->d_str_off); P.release(b->d_str_start); P.release(b->d_band_cells); P.release(b->d_info); P.release(b->d_band_qs); P.release(b->d_band_rs);
    for (auto e : b->ev) ev_put(b->ctx, e, true);
    for (auto e : b->ev_sync) ev_put(b->ctx, e, false);
    ev_put(b->ctx, b->ev_begin, true); ev_put(b->ctx, b->ev_end, true); ev_put(b->ctx, b->ev_h2d, false);
    delete b;
}

template <typename T>
static bool pool_alloc(dpx_ctx* ctx, T** p, size_t count) {
    *p = (T*)ctx->pool.alloc(std::max<size_t>(count, 1) * sizeof(T));
    if (!*p) { ctx->err = "device allocation failed"; return false; }
    return true;
}

// Upload, stage A (asynchronous): H2D of bytes [byte_lo, byte_hi) of the blob + the index, one device pass over
// the pairs (prep_kernel: alphabet, extrema, sizes, validity) and its 80-byte read-back into the lane's pinned slot.
static int batch_begin(dpx_ctx* ctx, cudaStream_t st, int lane, const char* sequences, long long byte_lo, long long byte_hi,
                       const dpx_seq_pair* pairs, size_t n_pairs, dpx_batch** out) {
    dpx_batch* b = new dpx_batch();
    b->ctx = ctx; b->stream = st; b->lane = lane; b->n_pairs = n_pairs; b->byte_lo = byte_lo; b->byte_hi = byte_hi;
    auto fail = [&](int s) { cudaStreamSynchronize(st); batch_release(b); return s; };
    const size_t nb = (size_t)(byte_hi - byte_lo);
    if (!pool_alloc(ctx, &b->d_blob_alloc, nb + 16) || !pool_alloc(ctx, &b->d_pairs, n_pairs) ||
        !pool_alloc(ctx, &b->d_scores, n_pairs) || !pool_alloc(ctx, &b->d_end_rc, 2 * n_pairs)) return fail(DPX_ERR_NOMEM);
    b->d_blob = b->d_blob_alloc - byte_lo;
    CUB_(ev_get(ctx, &b->ev_begin, true)); CUB_(ev_get(ctx, &b->ev_end, true));
    if (n_pairs == 0) { *out = b; return DPX_OK; }
    // inputs cross PCIe on the context's single copy stream (chunks arrive in issue order); the lane waits on the event
    CUB_(ev_get(ctx, &b->ev_h2d, false));
    if (nb) CUB_(cudaMemcpyAsync(b->d_blob_alloc, sequences + byte_lo, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUB_(cudaMemcpyAsync(b->d_pairs, pairs, n_pairs * sizeof(dpx_seq_pair), cudaMemcpyHostToDevice, ctx->copy_stream));
    CUB_(cudaEventRecord(b->ev_h2d, ctx->copy_stream));
    CUB_(cudaStreamWaitEvent(st, b->ev_h2d, 0));
    if (!pool_alloc(ctx, &b->d_info, 1) || !pool_alloc(ctx, &b->d_pk_off, n_pairs + 1) || !pool_alloc(ctx, &b->d_str_len, n_pairs + 1)) return fail(DPX_ERR_NOMEM);
    CUB_(cudaMemsetAsync(b->d_info, 0, sizeof(BatchInfo), st));
    const int pblocks = (int)std::min<size_t>((n_pairs + 7) / 8, (size_t)ctx->sm_count * 8);
    prep_kernel<<<pblocks, 256, 0, st>>>(b->d_blob, byte_lo, byte_hi, b->d_pairs, (int)n_pairs, b->d_info, b->d_pk_off, b->d_str_len);
    CUB_(cudaGetLastError());
    CUB_(cudaMemcpyAsync(ctx->h_info[lane], b->d_info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, st));
    *out = b;
    return DPX_OK;
}

// Upload, stage B: wait for stage A, read the batch facts, rank the alphabet and 2-bit pack when it has <= 4 symbols.
// On failure the batch is released.
static int batch_finish(dpx_batch* b) {
    dpx_ctx* ctx = b->ctx;
    cudaStream_t st = b->stream;
    const size_t n_pairs = b->n_pairs;
    auto fail = [&](int s) { cudaStreamSynchronize(st); batch_release(b); return s; };
    if (n_pairs == 0) return DPX_OK;
    CUB_(cudaStreamSynchronize(st));
    ctx->pool.release(b->d_info); b->d_info = nullptr;
    b->info = *ctx->h_info[b->lane];
    if (b->info.invalid) { ctx->err = "a seqPair entry points outside the sequence blob"; return fail(DPX_ERR_INVALID); }
    b->max_r = b->info.max_r; b->max_q = b->info.max_q;
    b->min_r = 0x7fffffff - b->info.min_r_inv; b->min_q = 0x7fffffff - b->info.min_q_inv;
    b->uniform = (b->max_r == b->min_r && b->max_q == b->min_q);
    PackLut lut; int nsym = 0;
    memset(lut.code, 0xFF, sizeof(lut.code));          // 0xFF = symbol not present in this batch
    for (int c = 0; c < 256; ++c) if (b->info.present[c >> 5] >> (c & 31) & 1u) lut.code[c] = (uint8_t)(nsym++ & 0x7f);
    b->n_symbols = nsym; b->lut = lut;
    const int pblocks = (int)std::min<size_t>((n_pairs + 7) / 8, (size_t)ctx->sm_count * 8);
    if (nsym <= 4) {
        if (!pool_alloc(ctx, &b->d_packed, (size_t)b->info.packed_words + 1)) return fail(DPX_ERR_NOMEM);
        const unsigned long long* off = nullptr;
        if (b->uniform) {
            b->pk_stride = (unsigned long long)((b->max_r + 15) >> 4) + (unsigned long long)((b->max_q + 15) >> 4);
            ctx->pool.release(b->d_pk_off); b->d_pk_off = nullptr;
        } else {
            // exclusive scan of the per-pair word counts, in place
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b->d_pk_off, b->d_pk_off, (int)n_pairs, st);
            void* tmp = ctx->pool.alloc(tmp_bytes);
            if (!tmp) return fail(DPX_ERR_NOMEM);
            CUB_(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, b->d_pk_off, b->d_pk_off, (int)n_pairs, st));
            CUB_(cudaStreamSynchronize(st));
            ctx->pool.release(tmp);
            off = b->d_pk_off;
        }
        pack2_kernel<<<pblocks, 256, 0, st>>>(b->d_blob, b->d_pairs, (int)n_pairs, off, b->pk_stride, b->d_packed, lut, nullptr);
        CUB_(cudaGetLastError());
        b->packed2 = true;
    } else {
        ctx->pool.release(b->d_pk_off); b->d_pk_off = nullptr;
        if (nsym <= 8) {
            // 5..8 symbols: a byte-coded copy of the blob for the WIDE pair-wavefront kernels (pairwf.cuh)
            const size_t nb = (size_t)(b->byte_hi - b->byte_lo);
            if (!pool_alloc(ctx, &b->d_codes_alloc, nb + 16)) return fail(DPX_ERR_NOMEM);
            b->d_codes = b->d_codes_alloc - b->byte_lo;
            code_bytes_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(b->d_blob_alloc, (long long)nb, b->d_codes_alloc, lut);
            CUB_(cudaGetLastError());
        }
    }
    return DPX_OK;
}

// Upload without the device pass and without a host sync, for chunks whose facts are already known from the host's
// scan of the index (uniform lengths).  Part 1 queues the copies; part 2 (once the alphabet map of the call is known)
// queues the pack kernel, which raises *d_unknown when it meets a byte outside that map (the caller then redoes the
// call the slow way).
// `stride` > 0: the chunk's index is an arithmetic progression (pair k at pairs[0] + k * stride, same sizes), so it is rebuilt
// on the device instead of crossing PCIe (16 bytes per pair: 5 % of a 150 x 150 batch's input).
static int batch_known_copy(dpx_ctx* ctx, cudaStream_t st, int lane, const char* sequences, long long byte_lo, long long byte_hi,
                            const dpx_seq_pair* pairs, size_t n_pairs, int R, int Q, long long stride, dpx_batch** out) {
    dpx_batch* b = new dpx_batch();
    b->ctx = ctx; b->stream = st; b->lane = lane; b->n_pairs = n_pairs; b->byte_lo = byte_lo; b->byte_hi = byte_hi;
    auto fail = [&](int s) { cudaStreamSynchronize(st); batch_release(b); return s; };
    const size_t nb = (size_t)(byte_hi - byte_lo);
    b->max_r = b->min_r = R; b->max_q = b->min_q = Q; b->uniform = true;
    b->info.max_r = R; b->info.max_q = Q; b->info.cells = (unsigned long long)n_pairs * R * Q;
    b->pk_stride = (unsigned long long)((R + 15) >> 4) + (unsigned long long)((Q + 15) >> 4);
    b->info.packed_words = b->pk_stride * n_pairs;
    if (!pool_alloc(ctx, &b->d_blob_alloc, nb + 16) || !pool_alloc(ctx, &b->d_pairs, n_pairs) ||
        !pool_alloc(ctx, &b->d_scores, n_pairs) || !pool_alloc(ctx, &b->d_end_rc, 2 * n_pairs) ||
        !pool_alloc(ctx, &b->d_packed, (size_t)b->info.packed_words + 1)) return fail(DPX_ERR_NOMEM);
    b->d_blob = b->d_blob_alloc - byte_lo;
    CUB_(ev_get(ctx, &b->ev_begin, true)); CUB_(ev_get(ctx, &b->ev_end, true));
    CUB_(ev_get(ctx, &b->ev_h2d, false));
    if (nb) CUB_(cudaMemcpyAsync(b->d_blob_alloc, sequences + byte_lo, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (stride > 0) {
        regular_pairs_kernel<<<(int)((n_pairs + 255) / 256), 256, 0, st>>>(b->d_pairs, (int)n_pairs, pairs[0], (int)stride);
        CUB_(cudaGetLastError());
    } else {
        CUB_(cudaMemcpyAsync(b->d_pairs, pairs, n_pairs * sizeof(dpx_seq_pair), cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CUB_(cudaEventRecord(b->ev_h2d, ctx->copy_stream));
    CUB_(cudaStreamWaitEvent(st, b->ev_h2d, 0));
    *out = b;
    return DPX_OK;
}

static int batch_known_pack(dpx_batch* b, const PackLut& lut, int nsym, int* d_unknown) {
    dpx_ctx* ctx = b->ctx;
    b->n_symbols = nsym; b->lut = lut;
    const int pblocks = (int)std::min<size_t>((b->n_pairs + 7) / 8, (size_t)ctx->sm_count * 8);
    pack2_kernel<<<pblocks, 256, 0, b->stream>>>(b->d_blob, b->d_pairs, (int)b->n_pairs, nullptr, b->pk_stride, b->d_packed, lut, d_unknown);
    CU(cudaGetLastError());
    b->packed2 = true;
    return DPX_OK;
}

static int batch_create(dpx_ctx* ctx, cudaStream_t st, int lane, const char* sequences, long long byte_lo, long long byte_hi,
                        const dpx_seq_pair* pairs, size_t n_pairs, dpx_batch** out) {
    dpx_batch* b = nullptr;
    int s = batch_begin(ctx, st, lane, sequences, byte_lo, byte_hi, pairs, n_pairs, &b);
    if (s) return s;
    s = batch_finish(b);
    if (s) return s;
    *out = b;
    return DPX_OK;
}
// Upload of pairs [p0, p0 + n) of a registered input from its host-side 2-bit sidecar (host_pack.h): the packed words cross
// PCIe as they are (0.25 B per base), plus 8 B per pair of sizes / word offsets when the lengths are ragged and nothing at
// all when they are uniform; no device pass over the bytes, no host sync.  Returns DPX_ERR_UNSUPPORTED when the chunk does
// not fit the synthetic int32 byte layout (the caller then takes the raw-byte route).
static int batch_from_sidecar(dpx_ctx* ctx, cudaStream_t st, int lane, const dpxhost_pack::Sidecar* sc, size_t p0, size_t n, dpx_batch** out) {
    const size_t w0 = sc->woff[p0], w1 = sc->woff[p0 + n], nw = w1 - w0;
    if (16ull * (unsigned long long)nw + 16ull > 0x7fffffffull) return DPX_ERR_UNSUPPORTED;
    dpx_batch* b = new dpx_batch();
    b->ctx = ctx; b->stream = st; b->lane = lane; b->n_pairs = n; b->byte_lo = 0; b->byte_hi = (long long)(16ull * nw);
    b->from_sidecar = true; b->packed2 = true; b->n_symbols = sc->nsym;
    memcpy(b->lut.code, sc->code, 256);
    b->inv4 = (uint32_t)sc->inv[0] | (uint32_t)sc->inv[1] << 8 | (uint32_t)sc->inv[2] << 16 | (uint32_t)sc->inv[3] << 24;
    auto fail = [&](int s) { cudaStreamSynchronize(st); batch_release(b); return s; };
    // facts of the chunk: closed form for uniform lengths, one host loop over the sizes otherwise
    if (sc->uniform) {
        b->max_r = b->min_r = sc->R; b->max_q = b->min_q = sc->Q; b->uniform = true;
        b->info.cells = (unsigned long long)n * (unsigned long long)sc->R * (unsigned long long)sc->Q;
        b->info.sum_r = (unsigned long long)n * sc->R; b->info.sum_q = (unsigned long long)n * sc->Q;
        b->pk_stride = (unsigned long long)((sc->R + 15) >> 4) + (unsigned long long)((sc->Q + 15) >> 4);
    } else if (n == sc->n_pairs) {
        b->max_r = sc->max_r; b->max_q = sc->max_q; b->min_r = sc->min_r; b->min_q = sc->min_q;
        b->info.cells = sc->cells; b->info.sum_r = sc->sum_r; b->info.sum_q = sc->sum_q;
    } else {
        int maxr = 0, maxq = 0, minr = 0x7fffffff, minq = 0x7fffffff; unsigned long long cells = 0, sr = 0, sq = 0;
        for (size_t i = p0; i < p0 + n; ++i) {
            const int R = sc->small ? (int)(sc->sizes[i] & 0xffffu) : (int)sc->sizes[2 * i], Q = sc->small ? (int)(sc->sizes[i] >> 16) : (int)sc->sizes[2 * i + 1];
            maxr = std::max(maxr, R); maxq = std::max(maxq, Q); minr = std::min(minr, R); minq = std::min(minq, Q);
            cells += (unsigned long long)R * (unsigned long long)Q; sr += (unsigned long long)R; sq += (unsigned long long)Q;
        }
        b->max_r = maxr; b->max_q = maxq; b->min_r = minr; b->min_q = minq;
        b->info.cells = cells; b->info.sum_r = sr; b->info.sum_q = sq;
    }
    if (!sc->uniform) b->uniform = (b->max_r == b->min_r && b->max_q == b->min_q);
    if (b->uniform && !sc->uniform) b->pk_stride = (unsigned long long)((b->max_r + 15) >> 4) + (unsigned long long)((b->max_q + 15) >> 4);
    b->info.max_r = b->max_r; b->info.max_q = b->max_q; b->info.packed_words = nw;
    b->info.str_bytes = 3ull * (b->info.sum_r + b->info.sum_q + (unsigned long long)n);
    uint32_t *d_sizes = nullptr, *d_woff = nullptr;
    const bool ragged = !sc->uniform;
    const size_t sz_words = ragged ? n * (sc->small ? 1 : 2) : 0;
    if (!pool_alloc(ctx, &b->d_packed, nw + 4) || !pool_alloc(ctx, &b->d_pairs, n) || !pool_alloc(ctx, &b->d_scores, n) ||
        !pool_alloc(ctx, &b->d_end_rc, 2 * n) || !pool_alloc(ctx, &b->d_str_len, n + 1) ||
        (!b->uniform && !pool_alloc(ctx, &b->d_pk_off, n + 1)) ||
        (ragged && (!pool_alloc(ctx, &d_sizes, sz_words) || !pool_alloc(ctx, &d_woff, n + 1)))) { ctx->pool.release(d_sizes); ctx->pool.release(d_woff); return fail(DPX_ERR_NOMEM); }
    auto fail2 = [&](int s) { cudaStreamSynchronize(st); cudaStreamSynchronize(ctx->copy_stream); ctx->pool.release(d_sizes); ctx->pool.release(d_woff); batch_release(b); return s; };
#define CUS_(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return fail2(e__ == cudaErrorMemoryAllocation ? DPX_ERR_NOMEM : DPX_ERR_CUDA); } } while (0)
    CUS_(ev_get(ctx, &b->ev_begin, true)); CUS_(ev_get(ctx, &b->ev_end, true));
    CUS_(ev_get(ctx, &b->ev_h2d, false));
    if (nw) CUS_(cudaMemcpyAsync(b->d_packed, sc->words + w0, nw * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (ragged) {
        CUS_(cudaMemcpyAsync(d_sizes, sc->sizes + p0 * (sc->small ? 1 : 2), sz_words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
        CUS_(cudaMemcpyAsync(d_woff, sc->woff + p0, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CUS_(cudaEventRecord(b->ev_h2d, ctx->copy_stream));
    CUS_(cudaStreamWaitEvent(st, b->ev_h2d, 0));
    sidecar_expand_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>((int)n, sc->R, sc->Q, b->pk_stride, d_sizes, sc->small ? 1 : 0, d_woff,
                                                                 b->d_pairs, b->uniform ? nullptr : b->d_pk_off, b->d_str_len);
    CUS_(cudaGetLastError());
#undef CUS_
    // the staging arrays go back to the pool now: the pool hands memory out in stream order of THIS lane only when the next
    // user is queued behind the expand kernel, which holds for every allocation made for this batch or after it on this lane;
    // other lanes could grab them earlier, so they stay with the batch until it is released
    b->d_stage[0] = d_sizes; b->d_stage[1] = d_woff;
    b->h2d_bytes = nw * sizeof(uint32_t) + (ragged ? (sz_words + n + 1) * sizeof(uint32_t) : 0);
    *out = b;
    return DPX_OK;
}

// Raw bytes of a sidecar batch, materialised on the device from the packed words the first time a kernel needs them
// (alignment strings, the byte-compare wavefront kernels).
static int ensure_blob(dpx_batch* b) {
    if (b->d_blob_alloc || !b->from_sidecar || b->n_pairs == 0) return DPX_OK;
    dpx_ctx* ctx = b->ctx;
    const unsigned long long nw = b->info.packed_words;
    if (!pool_alloc(ctx, &b->d_blob_alloc, (size_t)(16ull * nw) + 64)) return DPX_ERR_NOMEM;
    b->d_blob = b->d_blob_alloc;
    if (nw) {
        const int blocks = (int)std::min<unsigned long long>((nw + 255) / 256, (unsigned long long)ctx->sm_count * 16);
        unpack2_kernel<<<blocks, 256, 0, b->stream>>>(b->d_packed, nw, reinterpret_cast<uint4*>(b->d_blob_alloc), b->inv4);
        CU(cudaGetLastError());
    }
    return DPX_OK;
}
#undef CUB_

#include "host_run.cuh"

#include "host_long.cuh"

extern "C" {

void dpx_batch_free(dpx_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->stream);
    // the chunked traceback pipeline also queues work on the auxiliary streams; a run that failed half-way never joined them
    if (b->used_aux) for (int l = 0; l < 3; ++l) cudaStreamSynchronize(b->ctx->aux_stream[l]);
    batch_release(b);
}

int dpx_batch_upload(dpx_ctx* ctx, const char* sequences, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs,
                     dpx_batch** out) {
    if (!ctx || !out || (!sequences && n_bytes) || (!pairs && n_pairs) || n_pairs > 0x7fffffffu) return DPX_ERR_INVALID;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    size_t p0 = 0;
    if (const dpxhost_pack::Sidecar* sc = (n_pairs && !ctx->opt.no_sidecar) ? dpxhost_pack::find(sequences, n_bytes, pairs, n_pairs, &p0) : nullptr) {
        const int s = batch_from_sidecar(ctx, ctx->stream, 0, sc, p0, n_pairs, out);
        if (s != DPX_ERR_UNSUPPORTED) return s;
    }
    return batch_create(ctx, ctx->stream, 0, sequences, 0, (long long)n_bytes, pairs, n_pairs, out);
}

// ---- file-level entry points: the parser runs on the device ------------------------------------------
static void fill_input_info(const dpx_batch* b, size_t n_bytes, dpx_input_info* info) {
    if (!info) return;
    dpx_input_info in{};
    in.numPairs = b->n_pairs; in.numBytes = n_bytes;
    if (b->n_pairs) {
        in.numCells = (size_t)b->info.cells;
        in.maxReferenceLength = (size_t)b->max_r; in.maxQueryLength = (size_t)b->max_q;
        in.minReferenceLength = (size_t)b->min_r; in.minQueryLength = (size_t)b->min_q;
        in.avgReferenceLength = (double)b->info.sum_r / (double)b->n_pairs; in.avgQueryLength = (double)b->info.sum_q / (double)b->n_pairs;
    } else { in.minReferenceLength = SIZE_MAX; in.minQueryLength = SIZE_MAX; }
    *info = in;
}

int dpx_batch_upload_image(dpx_ctx* ctx, const char* image, size_t n_bytes, dpx_batch** out, dpx_input_info* info) {
    if (!ctx || !out || (!image && n_bytes) || n_bytes >= 0x7fffffffull) return DPX_ERR_INVALID;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    dpx_batch* b = new dpx_batch();
    b->ctx = ctx; b->stream = st; b->lane = 0; b->byte_lo = 0; b->byte_hi = (long long)n_bytes;
    int* d_nl = nullptr; int* d_count = nullptr; void* tmp = nullptr;
    auto fail = [&](int s) { cudaStreamSynchronize(st); ctx->pool.release(d_nl); ctx->pool.release(d_count); ctx->pool.release(tmp); batch_release(b); return s; };
#define CUI_(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return fail(e__ == cudaErrorMemoryAllocation ? DPX_ERR_NOMEM : DPX_ERR_CUDA); } } while (0)
    if (!pool_alloc(ctx, &b->d_blob_alloc, n_bytes + 16)) return fail(DPX_ERR_NOMEM);
    b->d_blob = b->d_blob_alloc;
    CUI_(ev_get(ctx, &b->ev_begin, true)); CUI_(ev_get(ctx, &b->ev_end, true));
    int n_lines = 0;
    if (n_bytes) {
        CUI_(cudaMemcpyAsync(b->d_blob_alloc, image, n_bytes, cudaMemcpyHostToDevice, st));
        // pass 1 counts the newlines (so the position list is sized exactly), pass 2 lists them
        if (!pool_alloc(ctx, &d_count, 1)) return fail(DPX_ERR_NOMEM);
        CUI_(cudaMemsetAsync(d_count, 0, sizeof(int), st));
        count_newlines_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(b->d_blob_alloc, (long long)n_bytes, d_count);
        CUI_(cudaGetLastError());
        CUI_(cudaMemcpyAsync(&n_lines, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUI_(cudaStreamSynchronize(st));
        if (n_lines % 3 != 0) return fail(DPX_ERR_FORMAT);                // parseInput.cpp:38-41
        if (!pool_alloc(ctx, &d_nl, (size_t)n_lines + 2)) return fail(DPX_ERR_NOMEM);
        size_t tmp_bytes = 0;
        thrust::counting_iterator<int> idx(0);
        cub::DeviceSelect::If(nullptr, tmp_bytes, idx, d_nl, d_count, (int)n_bytes, IsNewline{b->d_blob_alloc}, st);
        tmp = ctx->pool.alloc(tmp_bytes);
        if (!tmp) return fail(DPX_ERR_NOMEM);
        CUI_(cub::DeviceSelect::If(tmp, tmp_bytes, idx, d_nl, d_count, (int)n_bytes, IsNewline{b->d_blob_alloc}, st));
    }
    const size_t n_pairs = std::min<size_t>((size_t)n_lines / 3, 10000000); // INPUT_CAP, parseInput.cpp:7,102-105
    b->n_pairs = n_pairs;
    if (!pool_alloc(ctx, &b->d_pairs, n_pairs) || !pool_alloc(ctx, &b->d_scores, n_pairs) || !pool_alloc(ctx, &b->d_end_rc, 2 * n_pairs)) return fail(DPX_ERR_NOMEM);
    if (n_pairs) {
        pairs_from_newlines_kernel<<<(int)((n_pairs + 255) / 256), 256, 0, st>>>(d_nl, (int)n_pairs, b->d_pairs);
        CUI_(cudaGetLastError());
        if (!pool_alloc(ctx, &b->d_info, 1) || !pool_alloc(ctx, &b->d_pk_off, n_pairs + 1) || !pool_alloc(ctx, &b->d_str_len, n_pairs + 1)) return fail(DPX_ERR_NOMEM);
        CUI_(cudaMemsetAsync(b->d_info, 0, sizeof(BatchInfo), st));
        const int pblocks = (int)std::min<size_t>((n_pairs + 7) / 8, (size_t)ctx->sm_count * 8);
        prep_kernel<<<pblocks, 256, 0, st>>>(b->d_blob, 0, (long long)n_bytes, b->d_pairs, (int)n_pairs, b->d_info, b->d_pk_off, b->d_str_len);
        CUI_(cudaGetLastError());
        CUI_(cudaMemcpyAsync(ctx->h_info[0], b->d_info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, st));
    }
#undef CUI_
    ctx->pool.release(d_count); ctx->pool.release(tmp); d_count = nullptr; tmp = nullptr;
    int* keep_nl = d_nl; d_nl = nullptr;
    int s = batch_finish(b);                     // waits for the stream, ranks the alphabet, packs; releases b on failure
    ctx->pool.release(keep_nl);
    if (s) return s;
    fill_input_info(b, n_bytes, info);
    *out = b;
    return DPX_OK;
}

int dpx_align_file_text(dpx_ctx* ctx, const dpx_params* params, const char* path, long long first_index,
                        char** text, size_t* text_bytes, dpx_input_info* info) {
    if (!ctx || !params || !path || !text || !text_bytes) return DPX_ERR_INVALID;
    *text = nullptr; *text_bytes = 0;
    FILE* f = fopen(path, "rb");
    if (!f) return DPX_ERR_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return DPX_ERR_IO; }
    const long sz = ftell(f);
    if (sz < 0) { fclose(f); return DPX_ERR_IO; }
    rewind(f);
    char* img = (char*)g_host.take((size_t)sz + 1);               // page-locked: the upload runs at PCIe speed
    const bool pinned = img != nullptr;
    if (!img) img = (char*)malloc((size_t)sz + 1);
    if (!img) { fclose(f); return DPX_ERR_NOMEM; }
    const size_t got = fread(img, 1, (size_t)sz, f);
    fclose(f);
    auto drop = [&]() { if (pinned) g_host.give_back(img); else free(img); };
    if (got != (size_t)sz) { drop(); return DPX_ERR_IO; }
    dpx_batch* b = nullptr;
    int st = dpx_batch_upload_image(ctx, img, got, &b, info);
    drop();
    if (st) return st;
    st = dpx_batch_run(b, params);
    if (!st) st = dpx_batch_fetch_text(b, first_index, text, text_bytes);
    dpx_batch_free(b);
    return st;
}

int dpx_batch_run(dpx_batch* b, const dpx_params* p) {
    if (!b || !p) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    CU(cudaSetDevice(ctx->device));
    return batch_run(b, p);
}

int dpx_batch_sync(dpx_batch* b) {
    if (!b) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(b->stream));
    if (b->ran && b->ev_begin) {
        float ms = 0; double fill = 0, bt = 0;
        for (size_t k = 0; k < b->ev_kind.size(); ++k) {
            CU(cudaEventElapsedTime(&ms, b->ev[2 * k], b->ev[2 * k + 1]));
            (b->ev_kind[k] == 0 ? fill : bt) += ms;
        }
        CU(cudaEventElapsedTime(&ms, b->ev_begin, b->ev_end));
        b->stats.fill_ms = fill; b->stats.backtrack_ms = bt; b->stats.total_ms = ms;
        if ((b->params.algo == DPX_ALGO_BSW || b->params.algo == DPX_ALGO_ABSW) && b->d_band_cells) {
            unsigned long long c = 0;
            CU(cudaMemcpy(&c, b->d_band_cells, sizeof(c), cudaMemcpyDeviceToHost));
            b->stats.cells = c;
        }
    }
    return DPX_OK;
}

int dpx_batch_stats(const dpx_batch* b, dpx_run_stats* out) {
    if (!b || !out) return DPX_ERR_INVALID;
    *out = b->stats;
    return DPX_OK;
}

int dpx_batch_fetch(dpx_batch* b, int32_t* scores, int32_t* end_rc, char** strings_blob, size_t** string_offsets) {
    if (!b || !b->ran) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    const size_t n = b->n_pairs;
    CU(cudaSetDevice(ctx->device));
    if (strings_blob) *strings_blob = nullptr;
    if (string_offsets) *string_offsets = nullptr;
    { int s = batch_fetch_async(b, scores, end_rc); if (s) return s; }
    const bool want_strings = (b->params.flags & DPX_OUT_STRINGS) && strings_blob && string_offsets;
    char* blob = nullptr; size_t* offs = nullptr;
    auto drop = [&]() { if (blob) dpx_free(blob); if (offs) dpx_free(offs); blob = nullptr; offs = nullptr; };
    if (want_strings) {
        // The slab keeps 3 fields of Q+R+1 bytes per pair; what goes back to the host is the compacted form (3 NUL-terminated
        // strings of the alignment's own length per pair, in pair order) plus its offset table: a scan of the lengths, one copy
        // kernel, and two D2H copies into page-locked host memory.
        unsigned long long total = 0;
        unsigned long long *d_len = nullptr, *d_coff = nullptr, *d_offs = nullptr; char* d_compact = nullptr; void* tmp = nullptr;
        auto release = [&]() { ctx->pool.release(d_len); ctx->pool.release(d_coff); ctx->pool.release(d_offs); ctx->pool.release(d_compact); ctx->pool.release(tmp); };
        if (n) {
            if (!pool_alloc(ctx, &d_len, n + 1) || !pool_alloc(ctx, &d_coff, n + 1) || !pool_alloc(ctx, &d_offs, 3 * n)) { release(); return DPX_ERR_NOMEM; }
            str_len_kernel<<<(int)((n + 1 + 255) / 256), 256, 0, b->stream>>>(b->d_pairs, (int)n, b->d_str_start, d_len);
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_coff, (int)n + 1, b->stream);
            tmp = ctx->pool.alloc(tmp_bytes);
            if (!tmp) { release(); return DPX_ERR_NOMEM; }
            cudaError_t e0 = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, d_coff, (int)n + 1, b->stream);
            cudaError_t e1 = cudaMemcpyAsync(&total, d_coff + n, sizeof(total), cudaMemcpyDeviceToHost, b->stream);
            cudaError_t e2 = cudaStreamSynchronize(b->stream);
            if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) { release(); ctx->err = "string compaction failed"; return DPX_ERR_CUDA; }
        }
        static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t must be 64-bit");
        blob = (char*)g_host.take(std::max<size_t>((size_t)total, 1));
        offs = (size_t*)g_host.take(std::max<size_t>(3 * n, 1) * sizeof(size_t));
        if (!blob) blob = (char*)malloc(std::max<size_t>((size_t)total, 1));
        if (!offs) offs = (size_t*)malloc(std::max<size_t>(3 * n, 1) * sizeof(size_t));
        if (!blob || !offs) { release(); drop(); return DPX_ERR_NOMEM; }
        if (n) {
            if (!pool_alloc(ctx, &d_compact, (size_t)total + 16)) { release(); drop(); return DPX_ERR_NOMEM; }
            str_compact_kernel<<<(int)std::min<size_t>((n + 7) / 8, (size_t)ctx->sm_count * 16), 256, 0, b->stream>>>(
                b->d_pairs, (int)n, b->d_strings, b->d_str_off, b->d_str_start, d_coff, d_compact, d_offs);
            cudaError_t e1 = cudaMemcpyAsync(offs, d_offs, 3 * n * sizeof(size_t), cudaMemcpyDeviceToHost, b->stream);
            cudaError_t e2 = cudaMemcpyAsync(blob, d_compact, (size_t)total, cudaMemcpyDeviceToHost, b->stream);
            cudaError_t e3 = cudaStreamSynchronize(b->stream);
            release();
            if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { drop(); ctx->err = "string download failed"; return DPX_ERR_CUDA; }
        }
    }
    int st = dpx_batch_sync(b);
    if (st) { drop(); return st; }
    if (want_strings) { *strings_blob = blob; *string_offsets = offs; }
    return DPX_OK;
}

// The reference's stdout blocks for the whole batch, formatted on the device: one scan of the block
// lengths, one write kernel, one D2H copy into a page-locked blob.
int dpx_batch_fetch_text(dpx_batch* b, long long first_index, char** text, size_t* text_bytes) {
    if (!b || !b->ran || !text || !text_bytes) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    const size_t n = b->n_pairs;
    CU(cudaSetDevice(ctx->device));
    *text = nullptr; *text_bytes = 0;
    const bool strings = (b->params.flags & DPX_OUT_STRINGS) != 0;
    unsigned long long total = 0;
    unsigned long long *d_len = nullptr, *d_off = nullptr; char* d_text = nullptr; void* tmp = nullptr;
    auto release = [&]() { ctx->pool.release(d_len); ctx->pool.release(d_off); ctx->pool.release(d_text); ctx->pool.release(tmp); };
    if (n) {
        if (!pool_alloc(ctx, &d_len, n + 1) || !pool_alloc(ctx, &d_off, n + 1)) { release(); return DPX_ERR_NOMEM; }
        text_len_kernel<<<(int)((n + 1 + 255) / 256), 256, 0, b->stream>>>(b->d_pairs, (int)n, first_index, b->d_scores, strings ? b->d_str_start : nullptr, d_len);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_off, (int)n + 1, b->stream);
        tmp = ctx->pool.alloc(tmp_bytes);
        if (!tmp) { release(); return DPX_ERR_NOMEM; }
        cudaError_t e0 = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, d_off, (int)n + 1, b->stream);
        cudaError_t e1 = cudaMemcpyAsync(&total, d_off + n, sizeof(total), cudaMemcpyDeviceToHost, b->stream);
        cudaError_t e2 = cudaStreamSynchronize(b->stream);
        if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) { release(); ctx->err = "text formatting failed"; return DPX_ERR_CUDA; }
    }
    char* host = (char*)g_host.take(std::max<size_t>((size_t)total + 1, 1));
    if (!host) host = (char*)malloc(std::max<size_t>((size_t)total + 1, 1));
    if (!host) { release(); return DPX_ERR_NOMEM; }
    if (n) {
        if (!pool_alloc(ctx, &d_text, (size_t)total + 16)) { release(); dpx_free(host); return DPX_ERR_NOMEM; }
        text_write_kernel<<<(int)std::min<size_t>((n + 7) / 8, (size_t)ctx->sm_count * 16), 256, 0, b->stream>>>(
            b->d_pairs, (int)n, first_index, b->d_scores, b->d_strings, b->d_str_off, strings ? b->d_str_start : nullptr, d_off, d_text);
        cudaError_t e1 = cudaMemcpyAsync(host, d_text, (size_t)total, cudaMemcpyDeviceToHost, b->stream);
        cudaError_t e2 = cudaStreamSynchronize(b->stream);
        release();
        if (e1 != cudaSuccess || e2 != cudaSuccess) { dpx_free(host); ctx->err = "text download failed"; return DPX_ERR_CUDA; }
    }
    host[total] = 0;
    int st = dpx_batch_sync(b);
    if (st) { dpx_free(host); return st; }
    *text = host; *text_bytes = (size_t)total;
    return DPX_OK;
}

int dpx_align_batch_text(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                         const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                         int32_t* scores, int32_t* end_row_col, char** text, size_t* text_bytes) {
    if (!ctx || !params || !text || !text_bytes || (!sequences && n_bytes) || (!pairs && n_pairs) || n_pairs > 0x7fffffffu) return DPX_ERR_INVALID;
    dpx_batch* b = nullptr;
    int st = dpx_batch_upload(ctx, sequences, n_bytes, pairs, n_pairs, &b);
    if (st) return st;
    st = dpx_batch_run(b, params);
    if (!st && (scores || end_row_col)) { st = batch_fetch_async(b, scores, end_row_col); }
    if (!st) st = dpx_batch_fetch_text(b, first_index, text, text_bytes);
    dpx_batch_free(b);
    return st;
}

// ---- strings of one chunk of the one-call pipeline, in two asynchronous halves -------------------------------------------------
// begin : lengths of the compacted strings + their scan + the chunk's total into the lane's pinned slot (needs a host wait later);
// finish: (total known) compact into a device buffer and copy it, with offsets already rebased to `base`, into the shared host blob.
static int batch_strings_begin(dpx_batch* b, unsigned long long* h_total, unsigned long long** d_coff_out) {
    dpx_ctx* ctx = b->ctx;
    const size_t n = b->n_pairs;
    unsigned long long *d_len = nullptr, *d_coff = nullptr;
    if (!pool_alloc(ctx, &d_len, n + 1) || !pool_alloc(ctx, &d_coff, n + 1)) { ctx->pool.release(d_len); ctx->pool.release(d_coff); return DPX_ERR_NOMEM; }
    b->d_scratch.push_back(d_len); b->d_scratch.push_back(d_coff);
    str_len_kernel<<<(int)((n + 1 + 255) / 256), 256, 0, b->stream>>>(b->d_pairs, (int)n, b->d_str_start, d_len);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_coff, (int)n + 1, b->stream);
    void* tmp = ctx->pool.alloc(tmp_bytes);
    if (!tmp) return DPX_ERR_NOMEM;
    b->d_scratch.push_back(tmp);
    CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, d_coff, (int)n + 1, b->stream));
    CU(cudaMemcpyAsync(h_total, d_coff + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
    *d_coff_out = d_coff;
    return DPX_OK;
}

static int batch_strings_finish(dpx_batch* b, const unsigned long long* d_coff, unsigned long long total, unsigned long long base,
                                char* host_blob, size_t* host_offs /* [3 * n] of this chunk */) {
    dpx_ctx* ctx = b->ctx;
    const size_t n = b->n_pairs;
    char* d_compact = nullptr; unsigned long long* d_offs = nullptr;
    if (!pool_alloc(ctx, &d_compact, (size_t)total + 16) || !pool_alloc(ctx, &d_offs, 3 * n)) { ctx->pool.release(d_compact); ctx->pool.release(d_offs); return DPX_ERR_NOMEM; }
    b->d_scratch.push_back(d_compact); b->d_scratch.push_back(d_offs);
    str_compact_kernel<<<(int)std::min<size_t>((n + 7) / 8, (size_t)ctx->sm_count * 16), 256, 0, b->stream>>>(
        b->d_pairs, (int)n, b->d_strings, b->d_str_off, b->d_str_start, d_coff, d_compact, d_offs, base);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_offs, d_offs, 3 * n * sizeof(size_t), cudaMemcpyDeviceToHost, b->stream));
    if (total) CU(cudaMemcpyAsync(host_blob + base, d_compact, (size_t)total, cudaMemcpyDeviceToHost, b->stream));
    return DPX_OK;
}

// Alignment strings for a large batch: the pairs are cut into chunks that run on the four lanes (streams) like the score chunks
// below; a chunk's fill overlaps the previous chunk's walk, and its compaction + download overlap the next chunks' kernels, so the
// PCIe time of the strings (the largest part of the serial route's overhead) disappears behind the kernels.  The host blob is one
// page-locked allocation sized by the upper bound 3 (Q + R + 1) per pair; chunks land back to back at their exact sizes.
static int align_batch_strings_pipelined(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                                         const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                                         char** strings_blob, size_t** string_offsets, size_t nchunks) {
    constexpr int NL = 4;
    cudaStream_t lanes[NL] = {ctx->stream, ctx->aux_stream[0], ctx->aux_stream[1], ctx->aux_stream[2]};
    size_t sc_first = 0;
    const dpxhost_pack::Sidecar* sc = ctx->opt.no_sidecar ? nullptr : dpxhost_pack::find(sequences, n_bytes, pairs, n_pairs, &sc_first);
    std::vector<size_t> bound(nchunks + 1);
    for (size_t c = 0; c <= nchunks; ++c) bound[c] = n_pairs * c / nchunks;
    for (size_t c = 0; sc && c < nchunks; ++c)
        if (16ull * (unsigned long long)(sc->woff[sc_first + bound[c + 1]] - sc->woff[sc_first + bound[c]]) + 16ull > 0x7fffffffull) sc = nullptr;
    // upper bound of the compacted strings
    unsigned long long cap = 0;
    if (sc && n_pairs == sc->n_pairs) cap = 3ull * (sc->sum_r + sc->sum_q + (unsigned long long)n_pairs);
    else for (size_t i = 0; i < n_pairs; ++i) cap += 3ull * ((unsigned long long)std::max(pairs[i].referenceSize, 0) + (unsigned long long)std::max(pairs[i].querySize, 0) + 1ull);
    char* blob = (char*)g_host.take(std::max<size_t>((size_t)cap, 1));
    size_t* offs = (size_t*)g_host.take(std::max<size_t>(3 * n_pairs, 1) * sizeof(size_t));
    unsigned long long* h_total = nullptr;
    auto drop = [&]() { if (blob) dpx_free(blob); if (offs) dpx_free(offs); if (h_total) cudaFreeHost(h_total); };
    if (!blob || !offs || cudaHostAlloc(&h_total, NL * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); drop(); return DPX_ERR_NOMEM; }
    dpx_batch* inflight[NL] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<dpx_batch*> chunk_batch(nchunks, nullptr);
    std::vector<unsigned long long*> chunk_coff(nchunks, nullptr);
    std::vector<cudaEvent_t> chunk_ev(nchunks, nullptr);
    unsigned long long base = 0;
    int status = DPX_OK;
    auto stage_a = [&](size_t c) -> int {                    // upload (previous tenant of the lane is finished and released first)
        const int lane = (int)(c % NL);
        if (inflight[lane]) { cudaStreamSynchronize(lanes[lane]); batch_release(inflight[lane]); inflight[lane] = nullptr; }
        const size_t p0 = bound[c], n = bound[c + 1] - p0;
        dpx_batch* b = nullptr;
        int s = DPX_ERR_UNSUPPORTED;
        if (sc) s = batch_from_sidecar(ctx, lanes[lane], lane, sc, sc_first + p0, n, &b);
        if (s == DPX_ERR_UNSUPPORTED) {
            long long lo = (long long)n_bytes, hi = 0;
            for (size_t i = p0; i < p0 + n; ++i) {
                const dpx_seq_pair& q = pairs[i];
                lo = std::min<long long>(lo, std::min(q.referenceIdx, q.queryIdx));
                hi = std::max<long long>(hi, std::max((long long)q.referenceIdx + q.referenceSize, (long long)q.queryIdx + q.querySize));
            }
            if (lo < 0 || hi > (long long)n_bytes || hi < lo) { ctx->err = "a seqPair entry points outside the sequence blob"; return DPX_ERR_INVALID; }
            s = batch_create(ctx, lanes[lane], lane, sequences, lo, hi, pairs + p0, n, &b);
        }
        if (s) return s;
        inflight[lane] = b; chunk_batch[c] = b;
        return DPX_OK;
    };
    auto stage_b = [&](size_t c) -> int {                    // kernels + the first half of the strings
        dpx_batch* b = chunk_batch[c];
        const size_t p0 = bound[c];
        int s = batch_run(b, params);
        if (!s) s = batch_fetch_async(b, scores + p0, end_row_col ? end_row_col + 2 * p0 : nullptr);
        if (!s) s = batch_strings_begin(b, h_total + (c % NL), &chunk_coff[c]);
        if (!s) { cudaEvent_t e; if (ev_get(ctx, &e, false) != cudaSuccess || cudaEventRecord(e, b->stream) != cudaSuccess) s = DPX_ERR_CUDA; else chunk_ev[c] = e; }
        return s;
    };
    auto stage_c = [&](size_t c) -> int {                    // total known: compact + download at the running offset
        dpx_batch* b = chunk_batch[c];
        if (cudaEventSynchronize(chunk_ev[c]) != cudaSuccess) return DPX_ERR_CUDA;
        ev_put(ctx, chunk_ev[c], false); chunk_ev[c] = nullptr;
        const unsigned long long total = h_total[c % NL];
        if (base + total > cap) { ctx->err = "string blob overflow"; return DPX_ERR_CUDA; }
        const int s = batch_strings_finish(b, chunk_coff[c], total, base, blob, offs + 3 * bound[c]);
        base += total;
        return s;
    };
    status = stage_a(0);
    if (status == DPX_OK && nchunks > 1) status = stage_a(1);
    for (size_t c = 0; c < nchunks && status == DPX_OK; ++c) {
        status = stage_b(c);
        if (status == DPX_OK && c >= 1) status = stage_c(c - 1);           // (waits for chunk c-1 while chunk c's kernels are queued)
        if (status == DPX_OK && c + 2 < nchunks) status = stage_a(c + 2);
    }
    if (status == DPX_OK) status = stage_c(nchunks - 1);
    for (int lane = 0; lane < NL; ++lane) {
        cudaError_t e = cudaStreamSynchronize(lanes[lane]);
        if (e != cudaSuccess && status == DPX_OK) { ctx->err = std::string("stream sync: ") + cudaGetErrorString(e); status = DPX_ERR_CUDA; }
        if (inflight[lane]) batch_release(inflight[lane]);
    }
    for (auto e : chunk_ev) if (e) ev_put(ctx, e, false);
    cudaFreeHost(h_total); h_total = nullptr;
    if (status != DPX_OK) { drop(); return status; }
    *strings_blob = blob; *string_offsets = offs;
    return DPX_OK;
}

// One call, host buffers in and out.  Score / end-cell requests on large batches are cut into chunks of
// consecutive pairs that alternate between two streams, so the H2D copy of chunk k+1 (and the host's scan of
// its byte range) overlaps the kernels of chunk k; everything else takes the single-batch route.
static int align_batch_impl(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                            const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                            char** strings_blob, size_t** string_offsets);

static bool is_device_accessible_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int dpx_align_batch(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                    const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                    char** strings_blob, size_t** string_offsets) {
    if (!ctx || !params || !scores || (!sequences && n_bytes) || (!pairs && n_pairs) || n_pairs > 0x7fffffffu) return DPX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    // The chunk pipelines below copy results out asynchronously; a copy into PAGEABLE memory blocks the host thread and would
    // serialise the chunks (measured: config 3, 38 -> 51 ms).  Callers with ordinary malloc'ed result arrays (the reference's
    // driver style) therefore get page-locked bounce buffers and one memcpy at the end.
    if (n_pairs >= 65536 || ((params->flags & DPX_OUT_STRINGS) && n_pairs >= 16384)) {
        const bool bs = !is_device_accessible_host(scores), be = end_row_col && !is_device_accessible_host(end_row_col);
        if (bs || be) {
            int32_t* ts = bs ? (int32_t*)g_host.take(n_pairs * sizeof(int32_t)) : scores;
            int32_t* te = be ? (int32_t*)g_host.take(2 * n_pairs * sizeof(int32_t)) : end_row_col;
            if (ts && (te || !end_row_col)) {
                const int st = align_batch_impl(ctx, params, sequences, n_bytes, pairs, n_pairs, ts, te, strings_blob, string_offsets);
                if (st == DPX_OK) {
                    if (bs) memcpy(scores, ts, n_pairs * sizeof(int32_t));
                    if (be) memcpy(end_row_col, te, 2 * n_pairs * sizeof(int32_t));
                }
                if (bs) dpx_free(ts);
                if (be) dpx_free(te);
                return st;
            }
            if (bs && ts) dpx_free(ts);
            if (be && te) dpx_free(te);
        }
    }
    return align_batch_impl(ctx, params, sequences, n_bytes, pairs, n_pairs, scores, end_row_col, strings_blob, string_offsets);
}

static int align_batch_impl(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                            const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                            char** strings_blob, size_t** string_offsets) {
    const bool want_strings = (params->flags & DPX_OUT_STRINGS) != 0;
    const size_t min_chunk = 32768;
    // Registered input (parser output / dpx_register_input): every chunk uploads its slice of the host-side 2-bit sidecar --
    // a quarter of the bytes, no device pass, no alphabet handshake with chunk 0.
    size_t sc_first = 0;
    const dpxhost_pack::Sidecar* sc = (ctx->opt.no_sidecar || want_strings || n_pairs == 0) ? nullptr : dpxhost_pack::find(sequences, n_bytes, pairs, n_pairs, &sc_first);
    size_t nbase = want_strings ? 1 : std::min<size_t>((size_t)(sc ? ctx->chunks_packed : ctx->chunks), n_pairs / min_chunk);
    if (want_strings && strings_blob && string_offsets && !ctx->opt.serial_strings && n_pairs >= 16384) {
        // >= 2048 pairs per chunk, at most 8 chunks.  Measured (tools/e2e_strings.py): config 3 (100 k pairs) 43.5 ms serial, 38.1 ms in 8
        // chunks; config 4 (10 k long pairs) 18.5 ms serial, 20 ms in 4 chunks of 2500 walkers -- hence the floor of 16 k pairs.
        const size_t nstr = std::max<size_t>(1, std::min<size_t>((size_t)ctx->chunks_strings, n_pairs / 2048));
        *strings_blob = nullptr; *string_offsets = nullptr;
        return align_batch_strings_pipelined(ctx, params, sequences, n_bytes, pairs, n_pairs, scores, end_row_col, strings_blob, string_offsets, nstr);
    }
    if (nbase <= 1) {
        dpx_batch* b = nullptr;
        int st = dpx_batch_upload(ctx, sequences, n_bytes, pairs, n_pairs, &b);
        if (st) return st;
        st = dpx_batch_run(b, params);
        if (!st) st = dpx_batch_fetch(b, scores, end_row_col, strings_blob, string_offsets);
        dpx_batch_free(b);
        return st;
    }
    if (strings_blob) *strings_blob = nullptr;
    if (string_offsets) *string_offsets = nullptr;
    // Software pipeline over chunks on 4 lanes (streams); all input copies go through one copy stream in chunk order.
    //   chunk 0 (small): H2D + device pass (prep_kernel) + one host sync -> alphabet map of the call.
    //   later chunks whose index the host scan finds uniform: NO device pass and NO sync — H2D, pack (validating the
    //     alphabet), fill kernel and the D2H of the results are queued at once, so the GPU free-runs behind PCIe.
    //   other chunks: the two-stage form A (H2D + prep) / B (sync, pack, fill, D2H), queued two chunks ahead.
    // If a later chunk contains a byte outside chunk 0's alphabet the pack kernel raises a flag and the call is redone
    // through the single-batch route.
    constexpr int NL = 4;
    cudaStream_t lanes[NL] = {ctx->stream, ctx->aux_stream[0], ctx->aux_stream[1], ctx->aux_stream[2]};
    dpx_batch* inflight[NL] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<size_t> bound;
    {
        const size_t small = std::max<size_t>(n_pairs / 32, 8192);
        const bool edge = n_pairs >= 16 * small;
        const size_t lo = edge ? small : 0, hi = edge ? n_pairs - 2 * small : n_pairs;
        bound.push_back(0);
        for (size_t k = (edge ? 0 : 1); k < nbase; ++k) bound.push_back(lo + (hi - lo) * k / nbase);
        if (edge) { bound.push_back(hi); bound.push_back(hi + small); }
        bound.push_back(n_pairs);
    }
    const size_t nchunks = bound.size() - 1;
    std::vector<dpx_batch*> chunk_batch(nchunks, nullptr);
    int status = DPX_OK;
    const bool trace = ctx->opt.trace != 0;
    const auto t_call = std::chrono::steady_clock::now();
    auto now_us = [&]() { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_call).count(); };
    bool have_lut = false; PackLut lut; int nsym = 0;
    int* d_unknown = nullptr;
    if (!pool_alloc(ctx, &d_unknown, 1)) return DPX_ERR_NOMEM;
    CU(cudaMemsetAsync(d_unknown, 0, sizeof(int), ctx->copy_stream));

    for (size_t c = 0; sc && c < nchunks; ++c)
        if (16ull * (unsigned long long)(sc->woff[sc_first + bound[c + 1]] - sc->woff[sc_first + bound[c]]) + 16ull > 0x7fffffffull) sc = nullptr;

    std::vector<char> chunk_fast(nchunks, 0);
    // A(c): host scan of the chunk's index, buffers, input copies (and the device pass when the chunk is not uniform)
    auto stage_a = [&](size_t c) -> int {
        const size_t p0 = bound[c], p1 = bound[c + 1];
        const long long ta = now_us();
        if (sc) {
            const int lane = (int)(c % NL);
            if (inflight[lane]) { cudaStreamSynchronize(lanes[lane]); batch_release(inflight[lane]); inflight[lane] = nullptr; }
            dpx_batch* b = nullptr;
            const int s = batch_from_sidecar(ctx, lanes[lane], lane, sc, sc_first + p0, p1 - p0, &b);
            if (s) return s;
            inflight[lane] = b; chunk_batch[c] = b;
            if (trace) fprintf(stderr, "[dpx] A(%zu) sidecar pairs %zu bytes %zu  issue %lld..%lld us\n", c, p1 - p0, b->h2d_bytes, ta, now_us());
            return DPX_OK;
        }
        long long lo = (long long)n_bytes, hi = 0;
        int minr = 0x7fffffff, maxr = -1, minq = 0x7fffffff, maxq = -1;
        // is the index an arithmetic progression?  (fixed-length files: every record has the same size)
        const long long stride0 = p1 - p0 > 1 ? (long long)pairs[p0 + 1].referenceIdx - pairs[p0].referenceIdx : 1;
        const long long qoff0 = (long long)pairs[p0].queryIdx - pairs[p0].referenceIdx;
        bool regular = stride0 > 0 && stride0 < 0x7fffffff;
        for (size_t i = p0; i < p1; ++i) {
            const dpx_seq_pair& q = pairs[i];
            regular = regular && (long long)q.referenceIdx == (long long)pairs[p0].referenceIdx + (long long)(i - p0) * stride0 &&
                      (long long)q.queryIdx - q.referenceIdx == qoff0;
            const long long a0 = std::min(q.referenceIdx, q.queryIdx);
            const long long a1 = std::max((long long)q.referenceIdx + q.referenceSize, (long long)q.queryIdx + q.querySize);
            lo = std::min(lo, a0); hi = std::max(hi, a1);
            minr = std::min(minr, q.referenceSize); maxr = std::max(maxr, q.referenceSize);
            minq = std::min(minq, q.querySize); maxq = std::max(maxq, q.querySize);
        }
        if (lo < 0 || hi > (long long)n_bytes || hi < lo || minr < 0 || minq < 0) { ctx->err = "a seqPair entry points outside the sequence blob"; return DPX_ERR_INVALID; }
        const int lane = (int)(c % NL);
        if (inflight[lane]) {                       // the lane's previous chunk (c - NL): wait, then recycle its buffers
            cudaStreamSynchronize(lanes[lane]);
            batch_release(inflight[lane]); inflight[lane] = nullptr;
        }
        dpx_batch* b = nullptr;
        const bool fast = c > 0 && minr == maxr && minq == maxq;
        int s = fast ? batch_known_copy(ctx, lanes[lane], lane, sequences, lo, hi, pairs + p0, p1 - p0, maxr, maxq, regular ? stride0 : 0, &b)
                     : batch_begin(ctx, lanes[lane], lane, sequences, lo, hi, pairs + p0, p1 - p0, &b);
        if (s) return s;
        inflight[lane] = b; chunk_batch[c] = b; chunk_fast[c] = fast;
        if (trace) fprintf(stderr, "[dpx] A(%zu)%s pairs %zu bytes %lld  issue %lld..%lld us\n", c, fast ? " fast" : "", p1 - p0, hi - lo, ta, now_us());
        return DPX_OK;
    };
    // B(c): everything after the copies.  Fast chunks: pack (validating the alphabet) + fill + D2H, no sync.
    auto stage_b = [&](size_t c) -> int {
        const size_t p0 = bound[c];
        dpx_batch* b = chunk_batch[c];
        const long long tb0 = now_us();
        int s;
        if (sc) {
            s = batch_run(b, params);
            if (!s) s = batch_fetch_async(b, scores + p0, end_row_col ? end_row_col + 2 * p0 : nullptr);
            if (trace) fprintf(stderr, "[dpx] B(%zu) issued %lld..%lld us\n", c, tb0, now_us());
            return s;
        }
        if (chunk_fast[c] && have_lut && nsym <= 4) {
            s = batch_known_pack(b, lut, nsym, d_unknown);
        } else {
            if (chunk_fast[c]) {
                // alphabet too wide for the packed path: this chunk still needs its own facts -> device pass now
                const size_t n = b->n_pairs;
                if (!pool_alloc(ctx, &b->d_info, 1) || !pool_alloc(ctx, &b->d_pk_off, n + 1) || !pool_alloc(ctx, &b->d_str_len, n + 1)) return DPX_ERR_NOMEM;
                ctx->pool.release(b->d_packed); b->d_packed = nullptr;
                CU(cudaMemsetAsync(b->d_info, 0, sizeof(BatchInfo), b->stream));
                const int pblocks = (int)std::min<size_t>((n + 7) / 8, (size_t)ctx->sm_count * 8);
                prep_kernel<<<pblocks, 256, 0, b->stream>>>(b->d_blob, b->byte_lo, b->byte_hi, b->d_pairs, (int)n, b->d_info, b->d_pk_off, b->d_str_len);
                CU(cudaMemcpyAsync(ctx->h_info[b->lane], b->d_info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, b->stream));
            }
            s = batch_finish(b);
            if (s) { inflight[c % NL] = nullptr; return s; }          // batch_finish released it
            if (!have_lut) { have_lut = true; lut = b->lut; nsym = b->n_symbols; }
        }
        const long long tb1 = now_us();
        if (!s) s = batch_run(b, params);
        if (!s) s = batch_fetch_async(b, scores + p0, end_row_col ? end_row_col + 2 * p0 : nullptr);
        if (trace) fprintf(stderr, "[dpx] B(%zu) wait %lld..%lld us, issued by %lld us\n", c, tb0, tb1, now_us());
        return s;
    };
    // copies stay queued two chunks ahead of the chunk being processed; chunk 0 defines the alphabet map
    status = stage_a(0);
    if (status == DPX_OK && nchunks > 1) status = stage_a(1);
    for (size_t c = 0; c < nchunks && status == DPX_OK; ++c) {
        if (c + 2 < nchunks) status = stage_a(c + 2);
        if (status == DPX_OK) status = stage_b(c);
    }
    for (int lane = 0; lane < NL; ++lane) {
        cudaError_t e = cudaStreamSynchronize(lanes[lane]);
        if (e != cudaSuccess && status == DPX_OK) { ctx->err = std::string("stream sync: ") + cudaGetErrorString(e); status = DPX_ERR_CUDA; }
        if (inflight[lane]) batch_release(inflight[lane]);
    }
    int unknown = 0;
    if (status == DPX_OK && cudaMemcpy(&unknown, d_unknown, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) status = DPX_ERR_CUDA;
    ctx->pool.release(d_unknown);
    if (trace) fprintf(stderr, "[dpx] done %lld us%s\n", now_us(), unknown ? " (alphabet grew: redo)" : "");
    if (status == DPX_OK && unknown) {
        // a chunk used a symbol chunk 0 did not have: redo as one batch (device pass over everything)
        dpx_batch* b = nullptr;
        int st = dpx_batch_upload(ctx, sequences, n_bytes, pairs, n_pairs, &b);
        if (st) return st;
        st = dpx_batch_run(b, params);
        if (!st) st = dpx_batch_fetch(b, scores, end_row_col, nullptr, nullptr);
        dpx_batch_free(b);
        return st;
    }
    return status;
}

#include "host_long_abi.cuh"
#include "host_allmax.cuh"

}  // extern "C"
