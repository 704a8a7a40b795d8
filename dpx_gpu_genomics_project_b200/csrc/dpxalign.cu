// dpxalign.cu — C ABI (include/dpxalign.h) and host-side orchestration of libdpxalign.so.
// sm_100a only; no CPU fallback: every alignment entry point needs a CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"
#include "dpx_ops.cuh"
#include "wavefront.cuh"
#include "backtrack.cuh"
#include "pack.cuh"
#include "shortread.cuh"

using namespace dpx;

// ------------------------------------------------------------------------------------------------
struct dpx_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    // grow-only workspaces
    int32_t* boundary = nullptr; size_t boundary_ints = 0;
    unsigned int* counters = nullptr;              // 64 dynamic-work counters
    size_t tb_budget_bytes = (size_t)16 << 30;     // traceback chunk budget
};

struct dpx_batch {
    dpx_ctx* ctx = nullptr;
    size_t n_pairs = 0, n_bytes = 0;
    std::vector<dpx_seq_pair> h_pairs;
    int max_r = 0, max_q = 0, min_r = 0, min_q = 0;
    // device inputs
    uint8_t* d_blob = nullptr;
    dpx_seq_pair* d_pairs = nullptr;
    int32_t* d_order = nullptr;
    // 2-bit packed copy (pack.cuh); n_symbols = distinct sequence bytes in the batch
    uint32_t* d_packed = nullptr;
    unsigned long long* d_pk_off = nullptr;
    int n_symbols = 0;
    bool packed2 = false;
    std::vector<int32_t> h_order; bool order_ready = false; bool uniform = true;
    // device outputs
    int32_t* d_scores = nullptr;
    int32_t* d_end_rc = nullptr;
    uint32_t* d_tb = nullptr; size_t tb_words = 0;
    unsigned long long* d_tb_off = nullptr;
    char* d_strings = nullptr; size_t strings_bytes = 0;
    unsigned long long* d_str_off = nullptr;
    int32_t* d_str_start = nullptr;
    std::vector<unsigned long long> h_str_off;
    // run state
    bool ran = false; dpx_params params{};
    std::vector<cudaEvent_t> ev;     // pairs of (start, end) per kernel; kind in ev_kind
    std::vector<int> ev_kind;        // 0 fill, 1 backtrack
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    dpx_run_stats stats{};
};

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

int dpx_create(dpx_ctx** out, int device) {
    if (!out) return DPX_ERR_INVALID;
    *out = nullptr;
    int n = dpx_device_count();
    if (n <= 0) return DPX_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return DPX_ERR_INVALID;
    dpx_ctx* ctx = new dpx_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DPX_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DPX_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DPX_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    if (cudaMalloc(&ctx->counters, 64 * sizeof(unsigned int)) != cudaSuccess) { cudaStreamDestroy(ctx->own_stream); delete ctx; return DPX_ERR_NOMEM; }
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->tb_budget_bytes = std::min<size_t>((size_t)48 << 30, fr / 3);
    *out = ctx;
    return DPX_OK;
}

void dpx_destroy(dpx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->boundary) cudaFree(ctx->boundary);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* dpx_last_error(const dpx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int dpx_set_stream(dpx_ctx* ctx, void* s) {
    if (!ctx) return DPX_ERR_INVALID;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return DPX_OK;
}

void dpx_free(void* p) { free(p); }

// ---- parser (replaces c++/parseInput.cpp:9-119) -------------------------------------------------
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
    size_t lines = 0;
    for (size_t i = 0; i < got; ++i) lines += (buf[i] == '\n');
    if (lines % 3 != 0) { free(buf); return DPX_ERR_FORMAT; }      // parseInput.cpp:38-41
    size_t n = lines / 3;
    const size_t cap = 10000000;                                     // INPUT_CAP, parseInput.cpp:7,102-105
    dpx_seq_pair* idx = (dpx_seq_pair*)malloc(std::max<size_t>(n, 1) * sizeof(dpx_seq_pair));
    if (!idx) { free(buf); return DPX_ERR_NOMEM; }
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
    *pairs_out = idx; *seq_out = buf;
    if (info) *info = in;
    return DPX_OK;
}

// ---- DPX instruction evaluation -----------------------------------------------------------------
int dpx_dpx_eval(dpx_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, int n,
                 uint32_t* out, uint8_t* pred_hi, uint8_t* pred_lo) {
    if (!ctx || !a || !b || !c || !out || !pred_hi || !pred_lo || n < 0 || op < 0 || op >= OP_COUNT) return DPX_ERR_INVALID;
    if (n == 0) return DPX_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t *da, *db, *dc, *dout; uint8_t *dh, *dl;
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
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dout); cudaFree(dh); cudaFree(dl);
    return DPX_OK;
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
void dpx_batch_free(dpx_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->d_blob); cudaFree(b->d_pairs); cudaFree(b->d_order); cudaFree(b->d_packed); cudaFree(b->d_pk_off); cudaFree(b->d_scores); cudaFree(b->d_end_rc);
    cudaFree(b->d_tb); cudaFree(b->d_tb_off); cudaFree(b->d_strings); cudaFree(b->d_str_off); cudaFree(b->d_str_start);
    for (auto e : b->ev) cudaEventDestroy(e);
    if (b->ev_begin) cudaEventDestroy(b->ev_begin);
    if (b->ev_end) cudaEventDestroy(b->ev_end);
    delete b;
}

int dpx_batch_upload(dpx_ctx* ctx, const char* sequences, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs,
                     dpx_batch** out) {
    if (!ctx || !out || (!sequences && n_bytes) || (!pairs && n_pairs) || n_pairs > 0x7fffffffu) return DPX_ERR_INVALID;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    dpx_batch* b = new dpx_batch();
    b->ctx = ctx; b->n_pairs = n_pairs; b->n_bytes = n_bytes;
    b->h_pairs.assign(pairs, pairs + n_pairs);
    int maxr = 0, maxq = 0, minr = INT32_MAX, minq = INT32_MAX;
    for (size_t i = 0; i < n_pairs; ++i) {
        const dpx_seq_pair& p = pairs[i];
        if (p.referenceSize < 0 || p.querySize < 0 || p.referenceIdx < 0 || p.queryIdx < 0 ||
            (size_t)p.referenceIdx + (size_t)p.referenceSize > n_bytes || (size_t)p.queryIdx + (size_t)p.querySize > n_bytes) {
            delete b; return DPX_ERR_INVALID;
        }
        maxr = std::max(maxr, p.referenceSize); maxq = std::max(maxq, p.querySize);
        minr = std::min(minr, p.referenceSize); minq = std::min(minq, p.querySize);
    }
    b->max_r = maxr; b->max_q = maxq; b->min_r = n_pairs ? minr : 0; b->min_q = n_pairs ? minq : 0;
    auto fail = [&](int st) { dpx_batch_free(b); return st; };
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return fail(e__ == cudaErrorMemoryAllocation ? DPX_ERR_NOMEM : DPX_ERR_CUDA); } } while (0)
    CUB(cudaMalloc(&b->d_blob, std::max<size_t>(n_bytes, 16)));
    CUB(cudaMalloc(&b->d_pairs, std::max<size_t>(n_pairs, 1) * sizeof(dpx_seq_pair)));
    CUB(cudaMalloc(&b->d_scores, std::max<size_t>(n_pairs, 1) * sizeof(int32_t)));
    CUB(cudaMalloc(&b->d_end_rc, std::max<size_t>(n_pairs, 1) * 2 * sizeof(int32_t)));
    if (n_bytes) CUB(cudaMemcpyAsync(b->d_blob, sequences, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (n_pairs) CUB(cudaMemcpyAsync(b->d_pairs, pairs, n_pairs * sizeof(dpx_seq_pair), cudaMemcpyHostToDevice, ctx->stream));
    CUB(cudaEventCreate(&b->ev_begin)); CUB(cudaEventCreate(&b->ev_end));
    b->uniform = (b->max_q == b->min_q && b->max_r == b->min_r);
    if (n_pairs) {
        // alphabet scan -> rank codes -> 2-bit pack when <= 4 symbols (pack.cuh)
        uint32_t* d_present = nullptr;
        CUB(cudaMalloc(&d_present, 8 * sizeof(uint32_t)));
        CUB(cudaMemsetAsync(d_present, 0, 8 * sizeof(uint32_t), ctx->stream));
        const int pblocks = (int)std::min<size_t>((n_pairs + 7) / 8, (size_t)ctx->sm_count * 8);
        present_kernel<<<pblocks, 256, 0, ctx->stream>>>(b->d_blob, b->d_pairs, (int)n_pairs, d_present);
        uint32_t present[8];
        CUB(cudaMemcpyAsync(present, d_present, sizeof(present), cudaMemcpyDeviceToHost, ctx->stream));
        // word offsets of the packed copy while the scan runs
        std::vector<unsigned long long> pk_off(n_pairs + 1, 0);
        for (size_t i = 0; i < n_pairs; ++i)
            pk_off[i + 1] = pk_off[i] + (unsigned long long)((pairs[i].referenceSize + 15) >> 4) + (unsigned long long)((pairs[i].querySize + 15) >> 4);
        CUB(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_present);
        PackLut lut{}; int nsym = 0;
        for (int c = 0; c < 256; ++c) if (present[c >> 5] >> (c & 31) & 1u) lut.code[c] = (uint8_t)(nsym++ & 0xff);
        b->n_symbols = nsym;
        if (nsym <= 4) {
            CUB(cudaMalloc(&b->d_packed, std::max<size_t>(pk_off[n_pairs], 1) * sizeof(uint32_t)));
            CUB(cudaMalloc(&b->d_pk_off, (n_pairs + 1) * sizeof(unsigned long long)));
            CUB(cudaMemcpyAsync(b->d_pk_off, pk_off.data(), (n_pairs + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
            pack2_kernel<<<pblocks, 256, 0, ctx->stream>>>(b->d_blob, b->d_pairs, (int)n_pairs, b->d_pk_off, b->d_packed, lut);
            CUB(cudaGetLastError());
            CUB(cudaStreamSynchronize(ctx->stream));      // pk_off (host vector) must outlive the copy
            b->packed2 = true;
        }
    }
#undef CUB
    *out = b;
    return DPX_OK;
}

}  // extern "C"

template <int ALGO, bool TB, int K>
static int launch_wf(dpx_ctx* ctx, const WfArgs& a, int slots_wanted, int* blocks_out) {
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_fill_kernel<ALGO, TB, K>, 128, 0));
    if (per_sm < 1) per_sm = 1;
    int blocks = std::min(ctx->sm_count * per_sm, (slots_wanted + 3) / 4);
    if (blocks < 1) blocks = 1;
    *blocks_out = blocks;
    return DPX_OK;
}

template <int ALGO, bool TB, int K>
static void run_wf(const WfArgs& a, int blocks, cudaStream_t st) { wf_fill_kernel<ALGO, TB, K><<<blocks, 128, 0, st>>>(a); }

template <int ALGO, bool TB>
static int dispatch_wf_k(dpx_ctx* ctx, int K, const WfArgs& a, int blocks, bool query_only, int slots, int* blocks_out) {
    if (K == 4) { if (query_only) return launch_wf<ALGO, TB, 4>(ctx, a, slots, blocks_out); run_wf<ALGO, TB, 4>(a, blocks, ctx->stream); }
    else        { if (query_only) return launch_wf<ALGO, TB, 8>(ctx, a, slots, blocks_out); run_wf<ALGO, TB, 8>(a, blocks, ctx->stream); }
    return DPX_OK;
}

static int dispatch_wf(dpx_ctx* ctx, int algo, bool tb, int K, const WfArgs& a, int blocks, bool query_only, int slots, int* blocks_out) {
    switch (algo) {
        case DPX_ALGO_LNW: return tb ? dispatch_wf_k<DPX_ALGO_LNW, true>(ctx, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_LNW, false>(ctx, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_ANW: return tb ? dispatch_wf_k<DPX_ALGO_ANW, true>(ctx, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_ANW, false>(ctx, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_LSW: return tb ? dispatch_wf_k<DPX_ALGO_LSW, true>(ctx, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_LSW, false>(ctx, K, a, blocks, query_only, slots, blocks_out);
        case DPX_ALGO_BSW: return tb ? dispatch_wf_k<DPX_ALGO_BSW, true>(ctx, K, a, blocks, query_only, slots, blocks_out) : dispatch_wf_k<DPX_ALGO_BSW, false>(ctx, K, a, blocks, query_only, slots, blocks_out);
    }
    return DPX_ERR_INVALID;
}

template <int G, int K>
static int run_short(dpx_ctx* ctx, dpx_batch* b, const dpx_params* p, SrArgs a, bool track, bool xormode) {
    const int gpb = 128 / G;
    a.bnd_stride = b->max_r + G + 2;
    a.rsel_stride = (b->max_r + 2 * G + 2 + 1) & ~1;
    const size_t smem = (size_t)gpb * ((size_t)a.bnd_stride * 4 + (size_t)a.rsel_stride * 2);
    auto launch = [&](auto kern) -> int {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
        if (per_sm < 1) { ctx->err = "short-read kernel does not fit on an SM"; return DPX_ERR_RANGE; }
        const int warps_needed = (a.n_slots + (32 / G) - 1) / (32 / G);
        int blocks = std::min(ctx->sm_count * per_sm, (warps_needed + 3) / 4);
        if (blocks < 1) blocks = 1;
        kern<<<blocks, 128, smem, ctx->stream>>>(a);
        CU(cudaGetLastError());
        return DPX_OK;
    };
    if (track) return xormode ? launch(sr_lsw_kernel<G, K, true, true>) : launch(sr_lsw_kernel<G, K, true, false>);
    return xormode ? launch(sr_lsw_kernel<G, K, false, true>) : launch(sr_lsw_kernel<G, K, false, false>);
}

// Eligibility of the packed int16x2 short-read kernel (shortread.cuh).
static bool short_eligible(const dpx_batch* b, const dpx_params* p, int* B_out, bool* xormode, int* kbits_out) {
    if (p->algo != DPX_ALGO_LSW || (p->flags & DPX_OUT_STRINGS) || !b->packed2) return false;
    const int m = p->match, x = p->mismatch, g = p->gap_open;
    if (!(m > 0 && x < 0 && g < 0)) return false;                 // pads must stay strictly below real cells
    if (m - g > 127 || x - g < -128 || x - g > 127 || g < -4096) return false;
    if (b->max_r > 4096 || b->max_q > 65535) return false;
    const int B = std::max(2, -g);
    // position bits: (Hmax + B) << k < 32768; blocks of 2^k steps, at most 256 blocks per pass
    const long long top = (long long)m * std::min(b->max_r, b->max_q) + B;
    int k = 0;
    while (k < 8 && (top << (k + 1)) < 32768) ++k;
    if (k < 4) return false;
    if (((long long)b->max_r + 16) >> k >= 255) return false;
    *B_out = B; *xormode = (x - g < 0); *kbits_out = k;
    return true;
}

static uint64_t inband_cells(long long Q, long long R, long long W) {
    // sum over rows i=1..Q of |{j in [1,R] : |i-j| <= W}|
    uint64_t c = 0;
    for (long long i = 1; i <= Q; ++i) {
        long long lo = std::max(1LL, i - W), hi = std::min(R, i + W);
        if (hi >= lo) c += (uint64_t)(hi - lo + 1);
    }
    return c;
}

extern "C" {

int dpx_batch_run(dpx_batch* b, const dpx_params* p) {
    if (!b || !p) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    if (p->algo < DPX_ALGO_LNW || p->algo > DPX_ALGO_BSW) return DPX_ERR_INVALID;
    if (p->algo == DPX_ALGO_BSW && p->band < 0) return DPX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const size_t n = b->n_pairs;
    b->params = *p; b->ran = true;
    b->stats = dpx_run_stats{};
    for (auto e : b->ev) cudaEventDestroy(e);
    b->ev.clear(); b->ev_kind.clear();
    const bool want_strings = (p->flags & DPX_OUT_STRINGS) != 0;
    const int algo = p->algo;
    const int CB = (algo == DPX_ALGO_ANW) ? 4 : 2;
    const int K = (b->max_q <= 128) ? 4 : 8;
    int band = -1;
    if (algo == DPX_ALGO_BSW) band = std::min(p->band, std::max(b->max_q, b->max_r));

    // ---- statistics: cells ---------------------------------------------------------------------
    uint64_t cells = 0;
    for (size_t i = 0; i < n; ++i) {
        const dpx_seq_pair& pr = b->h_pairs[i];
        cells += (algo == DPX_ALGO_BSW) ? inband_cells(pr.querySize, pr.referenceSize, band)
                                        : (uint64_t)pr.querySize * (uint64_t)pr.referenceSize;
    }
    b->stats.cells = cells;
    b->stats.kernel_id = DPX_KERNEL_WAVEFRONT_S32;
    CU(cudaEventRecord(b->ev_begin, ctx->stream));
    if (n == 0) { CU(cudaEventRecord(b->ev_end, ctx->stream)); return DPX_OK; }

    // ---- schedule: longest pairs first when lengths differ (computed once per batch) ------------------
    const bool uniform = b->uniform;
    if (!uniform && !b->order_ready) {
        b->h_order.resize(n);
        std::iota(b->h_order.begin(), b->h_order.end(), 0);
        const auto& hp = b->h_pairs;
        // key: query length first (rows decide the short-read kernel's passes), then reference length
        std::stable_sort(b->h_order.begin(), b->h_order.end(), [&](int32_t x, int32_t y) {
            if (hp[x].querySize != hp[y].querySize) return hp[x].querySize > hp[y].querySize;
            return hp[x].referenceSize > hp[y].referenceSize; });
        if (!b->d_order) CU(cudaMalloc(&b->d_order, n * sizeof(int32_t)));
        CU(cudaMemcpyAsync(b->d_order, b->h_order.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        b->order_ready = true;
        CU(cudaEventRecord(b->ev_begin, ctx->stream));
    }
    const std::vector<int32_t>& order = b->h_order;
    auto pid_at = [&](size_t pos) -> size_t { return uniform ? pos : (size_t)order[pos]; };

    // ---- short-read path: packed int16x2 DPX kernel (score / end cell only) -------------------------
    {
        int B = 0, kbits = 0; bool xormode = false;
        if (short_eligible(b, p, &B, &xormode, &kbits)) {
            const int g = p->gap_open;
            SrArgs sa{};
            sa.packed = b->d_packed; sa.pk_off = b->d_pk_off; sa.pairs = b->d_pairs; sa.order = uniform ? nullptr : b->d_order;
            sa.n_pairs = (int)n; sa.n_slots = (int)((n + 1) / 2);
            const int ms = p->match - g, xs = p->mismatch - g;
            uint8_t tab[8];
            for (int k = 0; k < 8; ++k) tab[k] = (uint8_t)(int8_t)xs;
            tab[xormode ? 0 : 3] = (uint8_t)(int8_t)ms;
            sa.lut_lo = tab[0] | tab[1] << 8 | tab[2] << 16 | (uint32_t)tab[3] << 24;
            sa.lut_hi = tab[4] | tab[5] << 8 | tab[6] << 16 | (uint32_t)tab[7] << 24;
            auto pk = [](int v) { return (uint32_t)(v & 0xffff) | ((uint32_t)(v & 0xffff) << 16); };
            sa.one = 1u; sa.kbits = kbits; sa.kmul = 1u << kbits; sa.B = B; sa.B2 = pk(B); sa.Bg2 = pk(B + g);
            sa.G2 = (uint32_t)(g & 0xffff) | ((uint32_t)((g - 1) & 0xffff) << 16);
            sa.scores = b->d_scores; sa.end_rc = b->d_end_rc;
            sa.counter = ctx->counters;
            CU(cudaMemsetAsync(sa.counter, 0, sizeof(unsigned int), ctx->stream));
            const bool track = (p->flags & DPX_OUT_END_COORDS) != 0;
            cudaEvent_t s, e;
            CU(cudaEventCreate(&s)); CU(cudaEventCreate(&e));
            b->ev.push_back(s); b->ev.push_back(e); b->ev_kind.push_back(0);
            CU(cudaEventRecord(s, ctx->stream));
            int st;
            if (b->max_q <= 64) st = run_short<8, 8>(ctx, b, p, sa, track, xormode);
            else                st = run_short<8, 19>(ctx, b, p, sa, track, xormode);
            if (st) return st;
            CU(cudaEventRecord(e, ctx->stream));
            b->stats.kernel_launches = 1;
            b->stats.kernel_id = DPX_KERNEL_SHORT_S16X2;
            CU(cudaEventRecord(b->ev_end, ctx->stream));
            return DPX_OK;
        }
    }

    // ---- traceback + string slots; chunks of schedule positions under the traceback budget --------
    std::vector<unsigned long long> tb_off;
    std::vector<size_t> chunk_first{0};
    size_t tb_words_max = 0;
    if (want_strings) {
        tb_off.assign(n, 0);
        const size_t budget_words = std::max<size_t>(ctx->tb_budget_bytes / 4, 1);
        size_t cur = 0;
        for (size_t pos = 0; pos < n; ++pos) {
            const dpx_seq_pair& pr = b->h_pairs[pid_at(pos)];
            const size_t w = (size_t)WfGeom::make(K, CB, pr.querySize, pr.referenceSize, band).words();
            if (cur && cur + w > budget_words) { tb_words_max = std::max(tb_words_max, cur); chunk_first.push_back(pos); cur = 0; }
            tb_off[pid_at(pos)] = cur;
            cur += w;
        }
        tb_words_max = std::max(tb_words_max, cur);
        b->stats.traceback_bytes = 0;
        b->h_str_off.assign(n + 1, 0);
        for (size_t i = 0; i < n; ++i)
            b->h_str_off[i + 1] = b->h_str_off[i] + 3ull * ((unsigned long long)b->h_pairs[i].querySize + b->h_pairs[i].referenceSize + 1);
        if (b->tb_words < tb_words_max) {
            cudaFree(b->d_tb); b->d_tb = nullptr; b->tb_words = 0;
            CU(cudaMalloc(&b->d_tb, std::max<size_t>(tb_words_max, 1) * 4)); b->tb_words = tb_words_max;
        }
        if (!b->d_tb_off) CU(cudaMalloc(&b->d_tb_off, n * sizeof(unsigned long long)));
        if (!b->d_str_off) CU(cudaMalloc(&b->d_str_off, (n + 1) * sizeof(unsigned long long)));
        if (!b->d_str_start) CU(cudaMalloc(&b->d_str_start, n * sizeof(int32_t)));
        if (b->strings_bytes < b->h_str_off[n]) {
            cudaFree(b->d_strings); b->d_strings = nullptr; b->strings_bytes = 0;
            CU(cudaMalloc(&b->d_strings, std::max<size_t>(b->h_str_off[n], 1))); b->strings_bytes = b->h_str_off[n];
        }
        CU(cudaMemcpyAsync(b->d_tb_off, tb_off.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(b->d_str_off, b->h_str_off.data(), (n + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
        // total traceback bytes (all chunks)
        uint64_t tot = 0;
        for (size_t i = 0; i < n; ++i) tot += WfGeom::make(K, CB, b->h_pairs[i].querySize, b->h_pairs[i].referenceSize, band).words();
        b->stats.traceback_bytes = tot * 4;
    }
    chunk_first.push_back(n);

    // ---- boundary workspace ----------------------------------------------------------------------
    WfArgs a{};
    a.blob = b->d_blob; a.pairs = b->d_pairs; a.order = uniform ? nullptr : b->d_order;
    a.match = p->match; a.mismatch = p->mismatch; a.go = p->gap_open; a.ge = p->gap_extend; a.band = band;
    a.scores = b->d_scores; a.end_rc = b->d_end_rc;
    a.tb = want_strings ? b->d_tb : nullptr; a.tb_off = b->d_tb_off;
    a.rmax_p1 = b->max_r + 1;
    a.boundary_stride = 2LL * (b->max_r + 1);
    a.counter = ctx->counters;
    int blocks = 0;
    { int st = dispatch_wf(ctx, algo, want_strings, K, a, 0, true, (int)std::min<size_t>(n, 1u << 30), &blocks); if (st) return st; }
    const size_t need = (size_t)blocks * 4 * (size_t)a.boundary_stride;
    if (ctx->boundary_ints < need) {
        if (ctx->boundary) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->boundary); ctx->boundary = nullptr; ctx->boundary_ints = 0; }
        CU(cudaMalloc(&ctx->boundary, need * sizeof(int32_t))); ctx->boundary_ints = need;
    }
    a.boundary = ctx->boundary;

    auto add_event_pair = [&](int kind, cudaEvent_t* s, cudaEvent_t* e) -> int {
        CU(cudaEventCreate(s)); CU(cudaEventCreate(e));
        b->ev.push_back(*s); b->ev.push_back(*e); b->ev_kind.push_back(kind);
        return DPX_OK;
    };

    for (size_t c = 0; c + 1 < chunk_first.size(); ++c) {
        a.first = (int)chunk_first[c]; a.count = (int)(chunk_first[c + 1] - chunk_first[c]);
        a.counter = ctx->counters + (c % 64);
        CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), ctx->stream));
        cudaEvent_t s, e;
        { int st = add_event_pair(0, &s, &e); if (st) return st; }
        CU(cudaEventRecord(s, ctx->stream));
        dispatch_wf(ctx, algo, want_strings, K, a, blocks, false, 0, nullptr);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e, ctx->stream));
        b->stats.kernel_launches++;
        if (want_strings) {
            BtArgs t{};
            t.blob = b->d_blob; t.pairs = b->d_pairs; t.order = a.order; t.first = a.first; t.count = a.count;
            t.K = K; t.band = band; t.scores = b->d_scores; t.end_rc = b->d_end_rc; t.tb = b->d_tb; t.tb_off = b->d_tb_off;
            t.strings = b->d_strings; t.str_off = b->d_str_off; t.str_start = b->d_str_start;
            { int st = add_event_pair(1, &s, &e); if (st) return st; }
            CU(cudaEventRecord(s, ctx->stream));
            const int bt_blocks = (a.count + 127) / 128;
            switch (algo) {
                case DPX_ALGO_LNW: bt_walk_kernel<DPX_ALGO_LNW><<<bt_blocks, 128, 0, ctx->stream>>>(t); break;
                case DPX_ALGO_ANW: bt_walk_kernel<DPX_ALGO_ANW><<<bt_blocks, 128, 0, ctx->stream>>>(t); break;
                case DPX_ALGO_LSW: bt_walk_kernel<DPX_ALGO_LSW><<<bt_blocks, 128, 0, ctx->stream>>>(t); break;
                case DPX_ALGO_BSW: bt_walk_kernel<DPX_ALGO_BSW><<<bt_blocks, 128, 0, ctx->stream>>>(t); break;
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(e, ctx->stream));
            b->stats.kernel_launches++;
        }
    }
    CU(cudaEventRecord(b->ev_end, ctx->stream));
    return DPX_OK;
}

int dpx_batch_sync(dpx_batch* b) {
    if (!b) return DPX_ERR_INVALID;
    dpx_ctx* ctx = b->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (b->ran && b->ev_begin) {
        float ms = 0; double fill = 0, bt = 0;
        for (size_t k = 0; k < b->ev_kind.size(); ++k) {
            CU(cudaEventElapsedTime(&ms, b->ev[2 * k], b->ev[2 * k + 1]));
            (b->ev_kind[k] == 0 ? fill : bt) += ms;
        }
        CU(cudaEventElapsedTime(&ms, b->ev_begin, b->ev_end));
        b->stats.fill_ms = fill; b->stats.backtrack_ms = bt; b->stats.total_ms = ms;
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
    if (scores && n) CU(cudaMemcpyAsync(scores, b->d_scores, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (end_rc && n) CU(cudaMemcpyAsync(end_rc, b->d_end_rc, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    const bool want_strings = (b->params.flags & DPX_OUT_STRINGS) && strings_blob && string_offsets;
    char* blob = nullptr; size_t* offs = nullptr; std::vector<int32_t> starts;
    if (want_strings) {
        const size_t bytes = b->h_str_off.empty() ? 0 : (size_t)b->h_str_off[n];
        blob = (char*)malloc(std::max<size_t>(bytes, 1));
        offs = (size_t*)malloc(std::max<size_t>(3 * n, 1) * sizeof(size_t));
        if (!blob || !offs) { free(blob); free(offs); return DPX_ERR_NOMEM; }
        starts.resize(n);
        if (n) {
            CU(cudaMemcpyAsync(blob, b->d_strings, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(starts.data(), b->d_str_start, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    int st = dpx_batch_sync(b);
    if (st) { free(blob); free(offs); return st; }
    if (want_strings) {
        for (size_t i = 0; i < n; ++i) {
            const size_t F = (size_t)b->h_pairs[i].querySize + (size_t)b->h_pairs[i].referenceSize + 1;
            for (int k = 0; k < 3; ++k) offs[3 * i + k] = (size_t)b->h_str_off[i] + (size_t)k * F + (size_t)starts[i];
        }
        *strings_blob = blob; *string_offsets = offs;
    }
    return DPX_OK;
}

int dpx_align_batch(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                    const dpx_seq_pair* pairs, size_t n_pairs, int32_t* scores, int32_t* end_row_col,
                    char** strings_blob, size_t** string_offsets) {
    if (!ctx || !params || !scores) return DPX_ERR_INVALID;
    dpx_batch* b = nullptr;
    int st = dpx_batch_upload(ctx, sequences, n_bytes, pairs, n_pairs, &b);
    if (st) return st;
    st = dpx_batch_run(b, params);
    if (!st) st = dpx_batch_fetch(b, scores, end_row_col, strings_blob, string_offsets);
    dpx_batch_free(b);
    return st;
}

int dpx_align_long_pair(dpx_ctx* ctx, const dpx_params* params, const char* ref, size_t R, const char* qry, size_t Q,
                        int32_t* score, int64_t* end_row, int64_t* end_col) {
    if (!ctx || !params || !score || (!ref && R) || (!qry && Q)) return DPX_ERR_INVALID;
    if (R + Q + 2 > 0x7fffffffu) return DPX_ERR_RANGE;
    // Single-warp path for now: one pair through the batch engine (score + end cell only).
    std::vector<char> blob(R + Q + 2);
    if (R) memcpy(blob.data(), ref, R);
    blob[R] = 0;
    if (Q) memcpy(blob.data() + R + 1, qry, Q);
    blob[R + 1 + Q] = 0;
    dpx_seq_pair pr{0, (int32_t)R, (int32_t)(R + 1), (int32_t)Q};
    dpx_params p = *params; p.flags = DPX_OUT_SCORE | DPX_OUT_END_COORDS;
    int32_t rc[2] = {0, 0};
    int st = dpx_align_batch(ctx, &p, blob.data(), blob.size(), &pr, 1, score, rc, nullptr, nullptr);
    if (!st) { if (end_row) *end_row = rc[0]; if (end_col) *end_col = rc[1]; }
    return st;
}

}  // extern "C"
