// host_pack.cpp — see host_pack.h.  Plain host C++ (compiled by the host compiler through nvcc, linked into libdpxalign.so).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include "host_pack.h"

#include <cuda_runtime_api.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace dpxhost_pack {

namespace {

std::mutex g_mu;
std::unordered_map<const void*, Sidecar*> g_by_blob, g_by_pairs;

// The affinity mask the process had when the library was loaded: bind_thread_to_device narrows THIS mask, so a thread that was
// already bound next to one GPU (or inherited such a binding from its creator) can still be moved next to another.
cpu_set_t g_initial_mask;
bool g_have_initial = [] { return sched_getaffinity(0, sizeof(g_initial_mask), &g_initial_mask) == 0; }();

void* host_alloc(size_t bytes, bool* pinned) {
    bytes = std::max<size_t>(bytes, 64);
    void* p = nullptr;
    if (*pinned && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return p;
    cudaGetLastError();
    *pinned = false;
    return malloc(bytes);
}
void host_free(void* p, bool pinned) { if (!p) return; if (pinned) cudaFreeHost(p); else free(p); }

void destroy(Sidecar* s) {
    host_free(s->words, s->pinned); host_free(s->woff, s->pinned); host_free(s->sizes, s->pinned);
    delete s;
}

int n_threads(size_t bytes) {
    if (bytes < (size_t)4 << 20) return 1;
    unsigned hc = std::thread::hardware_concurrency();
    cpu_set_t set;                                             // respect a restricted affinity mask (containers, NUMA binding)
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = std::min<unsigned>(hc ? hc : 1, (unsigned)CPU_COUNT(&set));
    return (int)std::max(1u, std::min(hc ? hc : 1u, 32u));
}

template <typename F>
void parallel_ranges(size_t n, int threads, F&& fn) {          // fn(thread, begin, end) over [0, n)
    if (threads <= 1 || n < 2) { fn(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        const size_t b = n * (size_t)t / threads, e = n * (size_t)(t + 1) / threads;
        th.emplace_back([&fn, t, b, e] { fn(t, b, e); });
    }
    for (auto& x : th) x.join();
}

// 16 bases -> one word through the byte -> code table
inline uint32_t pack16_lut(const uint8_t* src, int n, const uint8_t* code) {
    uint32_t v = 0;
    for (int k = 0; k < n; ++k) v |= (uint32_t)(code[src[k]] & 3u) << (2 * k);
    return v;
}

// Fast path: the code is a fixed bit field of the byte ((c >> shift) & 3): eight bases per PEXT.
__attribute__((target("bmi2"))) void pack_seq_pext(const uint8_t* src, int len, uint32_t* out, unsigned long long mask, int shift, const uint8_t* code) {
    int k = 0, w = 0;
    for (; k + 16 <= len; k += 16, ++w) {
        unsigned long long a, b;
        memcpy(&a, src + k, 8); memcpy(&b, src + k + 8, 8);
        out[w] = (uint32_t)__builtin_ia32_pext_di(a, mask) | ((uint32_t)__builtin_ia32_pext_di(b, mask) << 16);
    }
    if (k < len) out[w] = pack16_lut(src + k, len - k, code);
    (void)shift;
}

void pack_seq_lut(const uint8_t* src, int len, uint32_t* out, const uint8_t* code) {
    int k = 0, w = 0;
    for (; k + 16 <= len; k += 16, ++w) out[w] = pack16_lut(src + k, 16, code);
    if (k < len) out[w] = pack16_lut(src + k, len - k, code);
}

// Fingerprint of the blob's size, head and tail: together with the size check of the first and last pair in find() it keeps a
// stale entry (memory released behind the library's back and reused for another input at the same address) from being used.
// Only memory the CALLER just handed in is read.
unsigned long long fingerprint_of(const char* blob, size_t n_bytes) {
    unsigned long long h = 1469598103934665603ull ^ (unsigned long long)n_bytes;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    const size_t k = std::min<size_t>(n_bytes, 64);
    mix(blob, k); mix(blob + n_bytes - k, k);
    return h;
}

}  // namespace

int register_input(const char* blob, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs) {
    if (!blob || !pairs || n_pairs == 0 || n_bytes == 0 || n_bytes > 0x7fffffffull || n_pairs > 0x7fffffffull) return DPX_OK;
    forget(blob);
    const uint8_t* B = reinterpret_cast<const uint8_t*>(blob);
    const int T = n_threads(n_bytes);
    // ---- pass 1: validity, facts, presence set ----------------------------------------------------------------
    struct Part { uint32_t present[8] = {0}; int max_r = 0, max_q = 0, min_r = 0x7fffffff, min_q = 0x7fffffff; bool bad = false, big = false;
                  unsigned long long cells = 0, sum_r = 0, sum_q = 0; };
    std::vector<Part> part((size_t)T);
    parallel_ranges(n_pairs, T, [&](int t, size_t b, size_t e) {
        Part& P = part[(size_t)t];
        bool seen[256] = {false};
        for (size_t i = b; i < e; ++i) {
            const dpx_seq_pair& q = pairs[i];
            if (q.referenceSize < 0 || q.querySize < 0 || q.referenceIdx < 0 || q.queryIdx < 0 ||
                (size_t)q.referenceIdx + (size_t)q.referenceSize > n_bytes || (size_t)q.queryIdx + (size_t)q.querySize > n_bytes) { P.bad = true; continue; }
            const uint8_t* r = B + q.referenceIdx; const uint8_t* s = B + q.queryIdx;
            for (int k = 0; k < q.referenceSize; ++k) seen[r[k]] = true;
            for (int k = 0; k < q.querySize; ++k) seen[s[k]] = true;
            P.max_r = std::max(P.max_r, q.referenceSize); P.max_q = std::max(P.max_q, q.querySize);
            P.min_r = std::min(P.min_r, q.referenceSize); P.min_q = std::min(P.min_q, q.querySize);
            P.cells += (unsigned long long)q.referenceSize * (unsigned long long)q.querySize;
            P.sum_r += (unsigned long long)q.referenceSize; P.sum_q += (unsigned long long)q.querySize;
            if (q.referenceSize > 65535 || q.querySize > 65535) P.big = true;
        }
        for (int c = 0; c < 256; ++c) if (seen[c]) P.present[c >> 5] |= 1u << (c & 31);
    });
    Part all;
    for (const Part& P : part) {
        for (int w = 0; w < 8; ++w) all.present[w] |= P.present[w];
        all.max_r = std::max(all.max_r, P.max_r); all.max_q = std::max(all.max_q, P.max_q);
        all.min_r = std::min(all.min_r, P.min_r); all.min_q = std::min(all.min_q, P.min_q);
        all.bad |= P.bad; all.big |= P.big; all.cells += P.cells; all.sum_r += P.sum_r; all.sum_q += P.sum_q;
    }
    if (all.bad) return DPX_ERR_INVALID;
    std::vector<int> sym;
    for (int c = 0; c < 256; ++c) if (all.present[c >> 5] >> (c & 31) & 1u) sym.push_back(c);
    if (sym.size() > 4) return DPX_OK;                         // a fifth symbol ('4' in the reference's data sets): raw-byte path

    Sidecar* s = new Sidecar();
    s->blob = blob; s->n_bytes = n_bytes; s->pairs = pairs; s->n_pairs = n_pairs; s->nsym = (int)sym.size();
    s->max_r = all.max_r; s->max_q = all.max_q; s->min_r = all.min_r; s->min_q = all.min_q;
    s->cells = all.cells; s->sum_r = all.sum_r; s->sum_q = all.sum_q;
    s->uniform = all.max_r == all.min_r && all.max_q == all.min_q; s->R = all.max_r; s->Q = all.max_q; s->small = !all.big;
    // code map: any injective map works (byte equality is all the aligners look at).  Prefer a bit field of the byte, which
    // PEXT extracts eight bases at a time: c & 3 separates '0'..'3'; (c >> 1) & 3 separates A C G T (either case).
    memset(s->code, 0xFF, sizeof(s->code));
    int shift = -1;
    for (int sh : {0, 1}) {
        bool used[4] = {false, false, false, false}, ok = true;
        for (int c : sym) { const int v = (c >> sh) & 3; if (used[v]) ok = false; used[v] = true; }
        if (ok) { shift = sh; break; }
    }
    s->inv[0] = s->inv[1] = s->inv[2] = s->inv[3] = sym.empty() ? (uint8_t)'0' : (uint8_t)sym[0];
    for (size_t k = 0; k < sym.size(); ++k) {
        const int v = shift >= 0 ? (sym[k] >> shift) & 3 : (int)k;
        s->code[sym[k]] = (uint8_t)v; s->inv[v] = (uint8_t)sym[k];
    }
    const bool pext_ok = shift >= 0 && __builtin_cpu_supports("bmi2");
    const unsigned long long mask = 0x0303030303030303ull << (shift > 0 ? shift : 0);

    // ---- word offsets ------------------------------------------------------------------------------------------
    s->pinned = true;
    s->woff = (uint32_t*)host_alloc(sizeof(uint32_t) * (n_pairs + 1), &s->pinned);
    if (!s->woff) { destroy(s); return DPX_ERR_NOMEM; }
    unsigned long long acc = 0;
    for (size_t i = 0; i < n_pairs; ++i) {
        s->woff[i] = (uint32_t)acc;
        acc += (unsigned long long)((pairs[i].referenceSize + 15) >> 4) + (unsigned long long)((pairs[i].querySize + 15) >> 4);
    }
    if (acc >= 0xffffffffull) { destroy(s); return DPX_OK; }
    s->woff[n_pairs] = (uint32_t)acc; s->n_words = (size_t)acc;
    bool pin2 = s->pinned;
    s->words = (uint32_t*)host_alloc(sizeof(uint32_t) * (s->n_words + 4), &pin2);
    if (pin2 != s->pinned || !s->words) {                       // keep one allocation kind for the whole sidecar
        if (s->words) host_free(s->words, pin2);
        s->words = nullptr; destroy(s); return DPX_OK;
    }
    if (!s->uniform) {
        bool pin3 = s->pinned;
        s->sizes = (uint32_t*)host_alloc(sizeof(uint32_t) * n_pairs * (s->small ? 1 : 2), &pin3);
        if (pin3 != s->pinned || !s->sizes) { if (s->sizes) host_free(s->sizes, pin3); s->sizes = nullptr; destroy(s); return DPX_OK; }
    }
    // ---- pass 2: pack -------------------------------------------------------------------------------------------
    parallel_ranges(n_pairs, T, [&](int, size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            const dpx_seq_pair& q = pairs[i];
            uint32_t* out = s->words + s->woff[i];
            const int rw = (q.referenceSize + 15) >> 4;
            if (pext_ok) { pack_seq_pext(B + q.referenceIdx, q.referenceSize, out, mask, shift, s->code); pack_seq_pext(B + q.queryIdx, q.querySize, out + rw, mask, shift, s->code); }
            else         { pack_seq_lut(B + q.referenceIdx, q.referenceSize, out, s->code); pack_seq_lut(B + q.queryIdx, q.querySize, out + rw, s->code); }
            if (s->sizes) {
                if (s->small) s->sizes[i] = (uint32_t)q.referenceSize | ((uint32_t)q.querySize << 16);
                else { s->sizes[2 * i] = (uint32_t)q.referenceSize; s->sizes[2 * i + 1] = (uint32_t)q.querySize; }
            }
        }
    });
    s->fingerprint = fingerprint_of(blob, n_bytes);
    std::lock_guard<std::mutex> g(g_mu);
    g_by_blob[blob] = s; g_by_pairs[pairs] = s;
    return DPX_OK;
}

const Sidecar* find(const char* blob, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs, size_t* first_pair) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_by_blob.find(blob);
    if (it == g_by_blob.end() || n_pairs == 0) return nullptr;
    const Sidecar* s = it->second;
    if (n_bytes != s->n_bytes || pairs < s->pairs || pairs > s->pairs + s->n_pairs) return nullptr;
    const size_t p0 = (size_t)(pairs - s->pairs);
    if (p0 + n_pairs > s->n_pairs) return nullptr;
    if (fingerprint_of(blob, n_bytes) != s->fingerprint) return nullptr;          // not the bytes that were packed
    auto same_sizes = [&](size_t i) {                                            // caller's pair i of its range vs. the packed record
        const size_t k = p0 + i;
        const int R = s->uniform ? s->R : (s->small ? (int)(s->sizes[k] & 0xffffu) : (int)s->sizes[2 * k]);
        const int Q = s->uniform ? s->Q : (s->small ? (int)(s->sizes[k] >> 16) : (int)s->sizes[2 * k + 1]);
        return pairs[i].referenceSize == R && pairs[i].querySize == Q;
    };
    if (!same_sizes(0) || !same_sizes(n_pairs - 1)) return nullptr;
    *first_pair = p0;
    return s;
}

const Sidecar* find_blob(const char* blob) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_by_blob.find(blob);
    return it == g_by_blob.end() ? nullptr : it->second;
}

void forget(const void* p) {
    Sidecar* s = nullptr;
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_by_blob.find(p);
        if (it != g_by_blob.end()) s = it->second;
        else { auto jt = g_by_pairs.find(p); if (jt != g_by_pairs.end()) s = jt->second; }
        if (!s) return;
        g_by_blob.erase(s->blob); g_by_pairs.erase(s->pairs);
    }
    destroy(s);
}

int bind_thread_to_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return 0;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return 0;
    cpu_set_t want, cur, both;
    CPU_ZERO(&want);
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; ++c) if (c >= 0) CPU_SET(c, &want);
    }
    if (g_have_initial) cur = g_initial_mask;
    else if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return 0;
    CPU_AND(&both, &want, &cur);
    const int n = CPU_COUNT(&both);
    if (n == 0) return 0;                                       // the container's cpuset does not reach that node: leave it alone
    if (pthread_setaffinity_np(pthread_self(), sizeof(both), &both) != 0) return 0;
    return n;
}

}  // namespace dpxhost_pack
