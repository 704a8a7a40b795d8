/* dpxalign.h — C ABI of libdpxalign.so, the B200 (sm_100a) pairwise-alignment engine.
 *
 * The reference (mickgordinier/DPX_GPU_Genomics_Project) has no FFI seam: its boundary is
 * the C++ class interface + parser + stdout format.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree).  The C++
 * shims in dpx_gpu_genomics_project_b200/host/ (SequenceAligner subclasses, parseInput,
 * drop-in main) and the Python ctypes mirror are built ONLY on this header.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; no stdout/stderr writes and
 * no exit() inside the library (the reference exit(1)s: c++/parseInput.cpp:14,28,40,58,63);
 * every function returns 0 on success or a negative dpx_status; results come back in pair
 * order; calls on one dpx_ctx must be serialised by the caller (one ctx per host thread).
 * There is NO CPU fallback: without a CUDA device dpx_create fails with DPX_ERR_NO_DEVICE.
 */
#ifndef DPXALIGN_H
#define DPXALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPX_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------- */
typedef enum {
    DPX_OK               = 0,
    DPX_ERR_INVALID      = -1,   /* bad argument (NULL, unknown algo, negative size, ...) */
    DPX_ERR_NO_DEVICE    = -2,   /* no usable CUDA device / driver */
    DPX_ERR_CUDA         = -3,   /* a CUDA runtime call or kernel failed; see dpx_last_error */
    DPX_ERR_NOMEM        = -4,   /* host or device allocation failed */
    DPX_ERR_IO           = -5,   /* dpx_parse_input / dpx_parse_fastx: cannot open / read the file */
    DPX_ERR_FORMAT       = -6,   /* dpx_parse_input: number of lines not a multiple of 3; dpx_parse_fastx: malformed or unpaired records */
    DPX_ERR_RANGE        = -7,   /* a sequence / score does not fit the selected kernel's range */
    DPX_ERR_UNSUPPORTED  = -8
} dpx_status;

/* ---- algorithms: one per reference SequenceAligner subclass ------------------------ */
typedef enum {
    DPX_ALGO_LNW = 0,   /* LinearNeedlemanWunsch   c++/LinearNeedlemanWunsch.{h,cpp}  */
    DPX_ALGO_ANW = 1,   /* AffineNeedlemanWunsch   c++/AffineNeedlemanWunsch.{h,cpp}  */
    DPX_ALGO_LSW = 2,   /* LinearSmithWaterman     c++/LinearSmithWaterman.{h,cpp}    */
    DPX_ALGO_BSW = 3,   /* BandedSmithWaterman     c++/BandedSmithWaterman.{h,cpp} (repaired semantics) */
    DPX_ALGO_ABSW = 4   /* affine banded Smith-Waterman: the variant the reference only names as a TODO
                           (python/LinearBandedSmithWaterman.py:8).  Local Gotoh on |i-j| <= band: D / I recurrences and the
                           tie -> GAP_OPEN rule of AffineNeedlemanWunsch (c++/AffineNeedlemanWunsch.cpp:185-213), ReLU,
                           UP > LEFT > DIAG, first-strict-max end cell and stop-at-zero walk of LinearSmithWaterman;
                           gap_open = 0 gives the linear BandedSmithWaterman with gap = gap_extend byte for byte.
                           band >= max(Q, R) (or any band larger than both lengths) = unbanded local Gotoh. */
} dpx_algo;

/* ---- output selection ------------------------------------------------------------ */
#define DPX_OUT_SCORE      0x1u   /* scores[]                      (always produced) */
#define DPX_OUT_END_COORDS 0x2u   /* end_row_col[]: SW first-max cell (1-based matrix indices); NW: (Q,R) */
#define DPX_OUT_STRINGS    0x4u   /* full traceback + the three alignment strings REF / REL / QRY */

/* Weights exactly as the reference constructors take them
 * (c++/LinearNeedlemanWunsch.h:42-47, c++/AffineNeedlemanWunsch.h:59-66,
 *  c++/LinearSmithWaterman.h:51-57, c++/BandedSmithWaterman.h:51-57).
 * Linear aligners use gap_open as their single gap weight (c++/main.cpp:238,244). */
typedef struct {
    int32_t  algo;          /* dpx_algo */
    int32_t  match;
    int32_t  mismatch;
    int32_t  gap_open;      /* linear: the gap weight */
    int32_t  gap_extend;    /* ANW, ABSW */
    int32_t  band;          /* BSW, ABSW: |i-j| <= band */
    uint32_t flags;         /* DPX_OUT_* */
} dpx_params;

/* Same layout as the reference's seqPair (c++/parseInput.h:22-29). */
typedef struct {
    int32_t referenceIdx;
    int32_t referenceSize;
    int32_t queryIdx;
    int32_t querySize;
} dpx_seq_pair;

/* Same fields as the reference's inputInfo (c++/parseInput.h:9-20). */
typedef struct {
    size_t numPairs;
    size_t numBytes;
    size_t numCells;
    size_t minReferenceLength;
    size_t minQueryLength;
    size_t maxReferenceLength;
    size_t maxQueryLength;
    double avgReferenceLength;
    double avgQueryLength;
} dpx_input_info;

typedef struct dpx_ctx   dpx_ctx;     /* one device, its streams and workspaces */
typedef struct dpx_batch dpx_batch;   /* a set of pairs resident in HBM + its results */

/* ---- library ------------------------------------------------------------------------ */
int         dpx_abi_version(void);
const char* dpx_strerror(int status);
/* Number of visible CUDA devices (0 without a driver/GPU; never fails). */
int         dpx_device_count(void);

/* ---- context -------------------------------------------------------------------------
 * dpx_create: device = CUDA ordinal.  Replaces nothing in the reference (it has no device
 * management: cuda/LNW/LinearNeedlemanWunschV19.cu:362-376 only prints the device count). */
int         dpx_create(dpx_ctx** out, int device);
void        dpx_destroy(dpx_ctx* ctx);
/* Last CUDA / internal error text of this ctx (valid until the next call on it). */
const char* dpx_last_error(const dpx_ctx* ctx);
/* Optional: run on an externally owned cudaStream_t (e.g. torch's current stream) so that the
 * caller's CUDA events bracket the kernels.  NULL restores the ctx's own stream. */
int         dpx_set_stream(dpx_ctx* ctx, void* cuda_stream);

/* Debug / test knobs (the library never reads the environment).  Unknown names return DPX_ERR_INVALID.
 *   "chunks" / "chunks_packed" (1..64)  middle chunks of the one-call pipeline (raw / sidecar input)   "tb_budget_bytes"  traceback slab budget
 *   "serial_chunks" 0/1      traceback chunks in series on one buffer              "trace" 0/1        chunk timeline on stderr
 *   "no_sidecar" 0/1         ignore the parser's packed copy, upload raw bytes     "no_shortread" / "no_pairwf" / "no_bandkernel" /
 *   "pairwf_int32" 0/1       kernel selection overrides (fall back to the next kernel family; still CUDA, never the CPU)
 *   "pairwf_k8" 0/1          Gotoh fill with 8 rows per lane where the library would take 16 (no traceback, queries above 256 rows)
 *   "long_k" (0,2,..32) "long_cap" "long_notable" "long_bt_tiles"   long-pair lane width / forced passes / byte-compare kernel / tiles per round */
int         dpx_set_option(dpx_ctx* ctx, const char* name, long long value);
/* Binds the calling host thread to the CPUs local to a CUDA device (its NUMA node) so that page-locked buffers allocated
 * afterwards sit next to that GPU's PCIe root.  Returns the number of CPUs in the new mask, 0 if nothing was changed. */
int         dpx_bind_host_to_device(int device);

/* ---- parser: replaces parseInput / cleanupParsedFile (c++/parseInput.cpp:9-119,140-143) --
 * Same file format and outputs (blob with '\n' -> '\0', seqPair index, inputInfo) but returns
 * DPX_ERR_IO / DPX_ERR_FORMAT instead of exit(1).  Both arrays are malloc'ed; release them with
 * dpx_free.  numCells is accumulated in 64-bit (the reference multiplies two ints, :100). */
int         dpx_parse_input(const char* path, dpx_seq_pair** pairs, char** sequences, dpx_input_info* info);
/* Same parser on a file image already in memory (copied; the caller keeps `image`). */
int         dpx_parse_image(const char* image, size_t n_bytes, dpx_seq_pair** pairs, char** sequences, dpx_input_info* info);
/* Releases anything the library handed out.  Page-locked result blobs are recycled through a process-wide cache of at most
 * 2 GiB; dpx_trim() (and the destruction of the last context) gives the cached blocks back to the driver. */
void        dpx_free(void* p);
void        dpx_trim(void);
/* Packed sidecar.  The parsers above also keep a 2-bit copy of the sequences in page-locked host memory (inputs with at most
 * four distinct symbols; the reference's timer likewise starts after parseInput, c++/main.cpp:157-164).  dpx_align_batch /
 * dpx_batch_upload / dpx_multi_align_batch recognise the blob pointer (with the index, or any sub-range of it) and move the
 * packed words over PCIe instead of the bytes: 4x fewer.  The blob and the index must therefore not be MODIFIED between parsing
 * and aligning (the reference's driver never does, c++/main.cpp:237-252); dpx_free of either drops the sidecar.
 * dpx_register_input does the same for a (blob, index) the caller built itself; dpx_unregister_input undoes it. */
int         dpx_register_input(const char* sequences, size_t n_bytes, const dpx_seq_pair* pairs, size_t n_pairs);
void        dpx_unregister_input(const char* sequences);
/* Introspection of a registered input's sidecar (tests, logging, byte accounting): 0 = `sequences` is not registered (more than
 * four symbols, or never parsed / registered); 1 = registered, ragged lengths (the upload adds 8 B per pair of sizes / offsets);
 * 2 = registered, uniform lengths (packed words only).  Pair p occupies words[word_offsets[p] .. word_offsets[p+1]): ceil(R/16)
 * reference words, then ceil(Q/16) query words, base k at bits 2*(k%16) of word k/16; code_to_byte[4] maps codes back to bytes.
 * Any output pointer may be NULL. */
int         dpx_input_sidecar(const char* sequences, size_t* n_pairs, size_t* n_words, const uint32_t** words, const uint32_t** word_offsets,
                              int* n_symbols, int* page_locked, unsigned char* code_to_byte);
/* FASTA / FASTQ front-end: the same blob + index from '>' / '@' records (multi-line sequences joined,
 * qualities skipped).  path_queries == NULL: the records of path_refs alternate reference, query; otherwise record k of
 * path_refs is the reference of pair k and record k of path_queries its query (DPX_ERR_FORMAT if the counts differ).
 * The blob is "ref\0qry\0ref\0qry\0..."; both arrays are malloc'ed (dpx_free). */
int         dpx_parse_fastx(const char* path_refs, const char* path_queries, dpx_seq_pair** pairs, char** sequences, dpx_input_info* info);

/* ---- one-call alignment: replaces the per-pair `Aligner a(ref, query, i, weights); a.align();`
 * loop of c++/main.cpp:237-252 for a whole batch.  Host buffers in, host buffers out
 * (H2D, kernels, D2H inside).
 *   sequences/n_bytes, pairs/n_pairs : the parseInput blob and index (borrowed)
 *   scores       [n_pairs]    caller-owned
 *   end_row_col  [2*n_pairs]  caller-owned or NULL
 *   strings_blob, string_offsets : NULL unless DPX_OUT_STRINGS; library-allocated (dpx_free):
 *       string k (0 REF, 1 REL, 2 QRY) of pair i is the NUL-terminated C string at
 *       (*strings_blob) + (*string_offsets)[3*i + k]  — the exact bytes the reference prints on
 *       the three lines after "<i> | <score>" (c++/LinearNeedlemanWunsch.cpp:207-213). */
int         dpx_align_batch(dpx_ctx* ctx, const dpx_params* params,
                            const char* sequences, size_t n_bytes,
                            const dpx_seq_pair* pairs, size_t n_pairs,
                            int32_t* scores, int32_t* end_row_col,
                            char** strings_blob, size_t** string_offsets);

/* ---- the same call on SEVERAL GPUs of one node from ONE host process (multi-GPU mode A): one worker thread
 * (bound to the device's NUMA node) + one dpx_ctx per device.  The pairs are cut into contiguous shards of equal cell count
 * (prefix sums of Q*R), so every GPU uploads one contiguous slice of the blob / sidecar; inside a shard the device scheduler
 * length-buckets as usual.  No collective: results are written straight into the caller's arrays, in pair order; strings are
 * stitched into one blob + offset table.  devices == NULL: devices 0 .. n_devices-1.
 * dpx_multi_shard_bounds: the shard boundaries the call would use (bounds[n_devices + 1]) -- for tests and logging. */
typedef struct dpx_multi dpx_multi;
int         dpx_create_multi(dpx_multi** out, const int* devices, int n_devices);
void        dpx_destroy_multi(dpx_multi* m);
int         dpx_multi_device_count(const dpx_multi* m);
const char* dpx_multi_last_error(const dpx_multi* m);
int         dpx_multi_set_option(dpx_multi* m, const char* name, long long value);     /* dpx_set_option on every device */
int         dpx_multi_align_batch(dpx_multi* m, const dpx_params* params,
                                  const char* sequences, size_t n_bytes,
                                  const dpx_seq_pair* pairs, size_t n_pairs,
                                  int32_t* scores, int32_t* end_row_col,
                                  char** strings_blob, size_t** string_offsets);
int         dpx_multi_align_batch_text(dpx_multi* m, const dpx_params* params,
                                       const char* sequences, size_t n_bytes,
                                       const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                                       int32_t* scores, int32_t* end_row_col, char** text, size_t* text_bytes);
int         dpx_multi_shard_bounds(const dpx_seq_pair* pairs, size_t n_pairs, int n_shards, size_t* bounds);

/* ---- staged form of the same call (what bench.py times stage by stage) ------------------
 * upload : H2D of blob + index, alphabet scan, 2-bit (4-bit escape) pack, length bucketing.
 * run    : fill kernels (+ GPU backtrack when DPX_OUT_STRINGS) on the ctx stream; asynchronous.
 * fetch  : waits for run, D2H of the selected outputs (same meaning as dpx_align_batch). */
int         dpx_batch_upload(dpx_ctx* ctx, const char* sequences, size_t n_bytes,
                             const dpx_seq_pair* pairs, size_t n_pairs, dpx_batch** out);
int         dpx_batch_run(dpx_batch* b, const dpx_params* params);
int         dpx_batch_sync(dpx_batch* b);
int         dpx_batch_fetch(dpx_batch* b, int32_t* scores, int32_t* end_row_col,
                            char** strings_blob, size_t** string_offsets);
void        dpx_batch_free(dpx_batch* b);

/* ---- formatted output: the exact bytes the reference prints for the batch, formatted on the GPU.
 * Per pair "<first_index + i> | <score>\n" (c++/LinearNeedlemanWunsch.cpp:207-209: printf("%d | ") + cout << score) followed,
 * when the run had DPX_OUT_STRINGS, by REF "\n" REL "\n" QRY "\n" (:210-213); a Smith-Waterman score of 0 gives three empty
 * lines (c++/LinearSmithWaterman.cpp:253-257).  *text is library-allocated (dpx_free), *text_bytes its length (a NUL follows).
 * dpx_align_batch_text = upload + run + fetch_text in one call; scores / end_row_col may be NULL. */
int         dpx_batch_fetch_text(dpx_batch* b, long long first_index, char** text, size_t* text_bytes);
int         dpx_align_batch_text(dpx_ctx* ctx, const dpx_params* params,
                                 const char* sequences, size_t n_bytes,
                                 const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                                 int32_t* scores, int32_t* end_row_col, char** text, size_t* text_bytes);

/* ---- parser on the GPU: replaces the two fread passes + byte loop of c++/parseInput.cpp:17-113.
 * dpx_batch_upload_image takes the FILE IMAGE (repeated "header\nREF\nQRY\n", newlines still in place): the newline scan, the
 * seqPair index, the alphabet scan and the 2-bit pack all run on the device; the batch then behaves like one made by
 * dpx_batch_upload (run / fetch / fetch_text).  DPX_ERR_FORMAT when the number of lines is not a multiple of 3
 * (parseInput.cpp:38-41); at most 10 000 000 pairs are taken (INPUT_CAP, :7,102-105).  info may be NULL.
 * dpx_align_file_text = read the file + upload_image + run + fetch_text: what the drop-in driver does for a whole input file. */
int         dpx_batch_upload_image(dpx_ctx* ctx, const char* file_image, size_t n_bytes, dpx_batch** out, dpx_input_info* info);
int         dpx_align_file_text(dpx_ctx* ctx, const dpx_params* params, const char* path, long long first_index,
                                char** text, size_t* text_bytes, dpx_input_info* info);

/* Statistics of the last dpx_batch_run on this batch (after dpx_batch_sync). */
typedef struct {
    double   fill_ms;          /* CUDA-event time of the fill kernel(s) */
    double   backtrack_ms;     /* CUDA-event time of the backtrack kernel(s) (0 if none) */
    double   total_ms;         /* first kernel -> last result byte resident on device */
    uint64_t cells;            /* sum of Q*R (BSW: in-band cells) */
    uint64_t traceback_bytes;  /* packed direction bytes written */
    uint32_t kernel_launches;  /* our kernels launched by the run */
    uint32_t kernel_id;        /* which fill kernel family ran (DPX_KERNEL_*) */
} dpx_run_stats;
int         dpx_batch_stats(const dpx_batch* b, dpx_run_stats* out);

#define DPX_KERNEL_WAVEFRONT_S32  1u  /* warp-per-pair anti-diagonal wavefront, int32 */
#define DPX_KERNEL_SHORT_S16X2    2u  /* thread-per-2-pairs, packed int16x2 DPX */
#define DPX_KERNEL_BAND_S32       3u  /* banded anti-diagonal, band mapped onto one warp */
#define DPX_KERNEL_TILED_S32      4u  /* multi-CTA tiled wavefront for one long pair */
#define DPX_KERNEL_PAIR_S32       6u  /* the same kernel family in int32, one pair per warp (scores beyond the int16 budget) */
#define DPX_KERNEL_PAIR_S16X2     5u  /* two pairs per warp, packed int16x2, directions in the low score bits (NW / Gotoh + traceback) */

/* ---- LinearSmithWaterman in the reference's BACKTRACK_ALL mode (c++/LinearSmithWaterman.h:9, .cpp:126-143,163-226):
 * one alignment per cell that holds the maximum, start cells queued bottom-right to top-left, finished alignments ordered by
 * their number of moves; text = the stdout blocks of the reference compiled with -DBACKTRACK_ALL ("<i> | <score>", then REF /
 * REL / QRY per alignment; score 0: three empty lines).  An inspection mode: whole int32 score matrices are kept on the device
 * (DPX_ERR_RANGE beyond 12 GiB).  text is library-allocated (dpx_free). */
int         dpx_align_batch_text_all(dpx_ctx* ctx, const dpx_params* params, const char* sequences, size_t n_bytes,
                                     const dpx_seq_pair* pairs, size_t n_pairs, long long first_index,
                                     char** text, size_t* text_bytes, long long* n_alignments);

/* ---- one very long pair, score + end cell only (BASELINE config 5) ----------------------
 * The reference cannot run this (8 B/cell full matrix). */
int         dpx_align_long_pair(dpx_ctx* ctx, const dpx_params* params,
                                const char* ref, size_t R, const char* qry, size_t Q,
                                int32_t* score, int64_t* end_row, int64_t* end_col);

/* ---- one very long pair WITH its alignment: the three lines LinearSmithWaterman::print_results
 * writes (c++/LinearSmithWaterman.cpp:259-285) for a pair whose direction matrix could never be stored.  The forward pass
 * keeps H on a grid of rows and columns (checkpoints); the walk of :160-226 then re-fills one tile at a time from its two
 * borders (csrc/longtrace.cuh).  Needs 4 B * R * Q * (1/512 + 1/1024) of device memory for the checkpoints (12 GB at 1 Mbp^2).
 *   lines      : library-allocated (dpx_free): REF, REL, QRY, each *line_len characters + NUL, back to back
 *   start_row/start_col : the cell where the walk stopped (H == 0); the alignment covers rows start_row+1 .. end_row
 *   stage_ms   : optional double[6]: forward pass (ms), walk rounds, tile fills + walk (ms), tiles walked, tile height, tile width */
int         dpx_align_long_pair_strings(dpx_ctx* ctx, const dpx_params* params,
                                        const char* ref, size_t R, const char* qry, size_t Q,
                                        int32_t* score, int64_t* end_row, int64_t* end_col,
                                        int64_t* start_row, int64_t* start_col,
                                        char** lines, size_t* line_len, double* stage_ms);

/* ---- one very long pair across SEVERAL GPUs (multi-GPU mode B): column stripes, one per GPU,
 * pipelined along the anti-diagonal; the right edge of stripe g streams into stripe g+1's inbox by NVLink P2P
 * stores.  One process per GPU: each rank creates its stripe, the ranks exchange the 64-byte IPC handles
 * (e.g. torch.distributed.all_gather_object), connect to their neighbours, reset, barrier, run, and reduce the
 * per-stripe (score, row, col) with the first-row-major tie rule.  The reference has no multi-GPU code at all.
 *   ref_stripe  : the R_local reference bases of this stripe = global columns col_offset+1 .. col_offset+R_local
 *   handle      : DPX_IPC_HANDLE_BYTES bytes */
#define DPX_IPC_HANDLE_BYTES 64
typedef struct dpx_stripe dpx_stripe;
int         dpx_stripe_create(dpx_ctx* ctx, const dpx_params* params, const char* ref_stripe, size_t R_local, size_t col_offset,
                              const char* qry, size_t Q, int stripe_index, int n_stripes, dpx_stripe** out);
int         dpx_stripe_export(dpx_stripe* s, void* handle);
int         dpx_stripe_connect(dpx_stripe* s, const void* prev_handle, const void* next_handle);
int         dpx_stripe_reset(dpx_stripe* s);      /* zero the exchange counters; all ranks must barrier before dpx_stripe_run */
int         dpx_stripe_run(dpx_stripe* s);        /* asynchronous */
int         dpx_stripe_result(dpx_stripe* s, int32_t* score, int64_t* end_row, int64_t* end_col, double* kernel_ms);
void        dpx_stripe_free(dpx_stripe* s);

/* ---- device self-test: the FakeDPX known-answer vectors (c++/testFakeDPX.cpp:10-113) run
 * against the real sm_100a DPX instructions.  Returns the number of failing vectors (0 = pass)
 * or a negative dpx_status. */
int         dpx_selftest_dpx(dpx_ctx* ctx);
/* Evaluates one DPX intrinsic (op = index into the list below) element-wise on the device:
 * out[i] = op(a[i], b[i], c[i]); pred_hi/pred_lo receive the predicate outputs of the __vib* forms
 * (pred_hi only for 32-bit forms).  Lets tests replay every FakeDPX vector on real hardware.
 * op order: vimax3 {s32,s16x2,u32,u16x2}, vimin3 {same}, vimax_relu {s32,s16x2}, vimin_relu {s32,s16x2},
 * vimax3_relu {s32,s16x2}, vimin3_relu {s32,s16x2}, vibmax {s32,u32}, vibmin {s32,u32},
 * vibmax {s16x2,u16x2}, vibmin {s16x2,u16x2}, viaddmax {s32,u32}, viaddmin {s32,u32},
 * viaddmax {s16x2,u16x2}, viaddmin {s16x2,u16x2}, viaddmax_relu s32, viaddmin_relu s32,
 * viaddmax_relu s16x2, viaddmin_relu s16x2   (36 ops; c++/FakeDPX.hpp). */
int         dpx_dpx_eval(dpx_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, int n,
                         uint32_t* out, uint8_t* pred_hi, uint8_t* pred_lo);

#ifdef __cplusplus
}
#endif
#endif /* DPXALIGN_H */
