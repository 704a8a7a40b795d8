// host_pack.h — host-side 2-bit "sidecar" of a parseInput blob (plain C++, no CUDA kernels; host_pack.cpp).
//
// The reference's aligners only compare bytes for equality (c++/LinearNeedlemanWunsch.cpp:108), so every kernel of this
// library works on 2-bit codes when the input has at most four symbols.  Shipping the raw 1-byte-per-base blob over PCIe
// just to shrink it 4x on the device made the one-call path PCIe-bound (round 1: 320 MB per 1M short-read pairs).  The
// parser (dpx_parse_input / dpx_parse_image) and dpx_register_input therefore pack ONCE on the host, multi-threaded, into
// page-locked memory; dpx_align_batch / dpx_batch_upload recognise a registered blob pointer and upload the packed words
// (plus 8 bytes per pair of sizes / offsets when the lengths are ragged, nothing when they are uniform) instead of the bytes.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/dpxalign.h"

namespace dpxhost_pack {

struct Sidecar {
    const char* blob = nullptr; size_t n_bytes = 0;
    const dpx_seq_pair* pairs = nullptr; size_t n_pairs = 0;
    int nsym = 0;                         // distinct sequence bytes; the packed form exists only when nsym <= 4
    uint8_t code[256];                    // byte -> 2-bit code (0xFF: byte not present)
    uint8_t inv[4];                       // code -> byte
    bool uniform = false; int R = 0, Q = 0;   // every pair has the same (referenceSize, querySize)
    bool small = false;                   // every size < 65536: sizes[] holds R | Q << 16 per pair, else (R, Q) as two words
    // page-locked when a CUDA device is present (plain malloc otherwise), all indexed by pair:
    uint32_t* words = nullptr; size_t n_words = 0;   // pair p: words[woff[p] .. woff[p+1]) = ceil(R/16) reference words, then ceil(Q/16) query words
    uint32_t* woff = nullptr;             // [n_pairs + 1] word offsets (n_words < 2^32 because the blob is < 2 GiB)
    uint32_t* sizes = nullptr;            // ragged only
    bool pinned = false;
    int max_r = 0, max_q = 0, min_r = 0, min_q = 0;
    unsigned long long cells = 0, sum_r = 0, sum_q = 0;
    unsigned long long fingerprint = 0;   // of the blob's size, head and tail (see find())
};

// Builds and registers the sidecar of (blob, pairs).  Returns DPX_OK also when the input cannot be packed (> 4 symbols):
// nothing is registered then and the raw-byte upload is used.  DPX_ERR_INVALID for an index that leaves the blob.
int register_input(const char* blob, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs);
// Looks (blob, [pairs, pairs + n_pairs)) up: the pairs may be any sub-range of the registered index.  Returns the sidecar and
// the position of pairs[0] in it, or nullptr.  The sidecar stays valid until the blob or the index is released.
const Sidecar* find(const char* blob, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs, size_t* first_pair);
const Sidecar* find_blob(const char* blob);
// Called by dpx_free for every pointer it frees: drops the sidecar whose blob or index this is.
void forget(const void* p);

// Binds the calling thread to the CPUs local to a CUDA device (sysfs local_cpulist of its PCI function), so that page-locked
// buffers it then allocates land on that NUMA node.  Returns the number of CPUs in the mask (0: unknown, nothing changed).
int bind_thread_to_device(int device);

}  // namespace dpxhost_pack
