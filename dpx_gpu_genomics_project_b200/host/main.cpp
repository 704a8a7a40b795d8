// main.cpp — drop-in driver: same command line and the same stdout bytes as the reference driver
// (c++/main.cpp:118-262), but the whole file goes to the GPU as ONE batch instead of one aligner object per pair.
//
//   main -pairs <file> [-match m] [-mismatch x] [-open g] [-extend e]          (positional, as the reference :133-150)
//   extra, after the reference's flags:  -algo LSW|LNW|ANW|BSW|ABSW   (the reference picks it with a #define, :22-24;
//                                                                default LSW = what the reference ships enabled)
//                                        -band W                 (BSW only, default 64)
//                                        -scores                 score (+ end cell) lines only, no alignment strings
//                                        -all                    LSW, every maximum cell walked: the output of the reference built with
//                                                                -DBACKTRACK_ALL (c++/LinearSmithWaterman.h:9)
//                                        -long                   LSW on pairs too long for a batch (Mbp): every pair of the file goes
//                                                                through dpx_align_long_pair_strings (checkpoints + tile walk), same blocks
//                                        -gpus N | -devices a,b,..  every GPU named aligns one contiguous shard of the file (dpx_create_multi:
//                                                                one host thread + context per device, results in pair order)
//                                        -fastx F [-fastx2 G]    pairs from FASTA / FASTQ records instead of -pairs (one file: records
//                                                                alternate reference, query; two files: paired by order)
// Output: "Parsing input file: F", "Pair # | Score", then per pair "<i> | <score>" REF REL QRY, then
// "Elapsed time (usec): N" and "Cleaning up" (:153,165,257,260).  Linear aligners take -open as their gap (:238,244).
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "GpuAligner.h"
#include "parseInput.h"

int main(int argc, char* argv[]) {
    if (argc < 2) {
        fprintf(stderr, "usage: main -pairs <InSeqFile> -match <matchWeight> -mismatch <mismatchWeight> -open <gapOpen> [-extend <gapExtend>] [-algo LSW|LNW|ANW|BSW] [-band W] [-scores]\n");
        exit(EXIT_FAILURE);
    }
    const char* pairFileName = nullptr;
    int matchWeight = 3, mismatchWeight = -1, gapOpenWeight = -4, gapExtendWeight = -1;   // defaults of main.cpp:128-132
    int algo = DPX_ALGO_LSW, band = 64; bool scores_only = false, long_pairs = false, all_maxima = false;
    const char* fastx = nullptr; const char* fastx2 = nullptr;
    std::vector<int> devices;
    for (int i = 1; i < argc; ++i) {
        const bool has = i + 1 < argc;
        if (!strcmp(argv[i], "-pairs") && has) pairFileName = argv[++i];
        else if (!strcmp(argv[i], "-match") && has) matchWeight = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-mismatch") && has) mismatchWeight = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-open") && has) gapOpenWeight = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-extend") && has) gapExtendWeight = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-band") && has) band = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-scores")) scores_only = true;
        else if (!strcmp(argv[i], "-long")) long_pairs = true;
        else if (!strcmp(argv[i], "-all")) all_maxima = true;
        else if (!strcmp(argv[i], "-gpus") && has) { const int n = atoi(argv[++i]); devices.clear(); for (int d = 0; d < n; ++d) devices.push_back(d); }
        else if (!strcmp(argv[i], "-devices") && has) { devices.clear(); for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t)); }
        else if (!strcmp(argv[i], "-fastx") && has) fastx = argv[++i];
        else if (!strcmp(argv[i], "-fastx2") && has) fastx2 = argv[++i];
        else if (!strcmp(argv[i], "-algo") && has) {
            const char* a = argv[++i];
            if (!strcmp(a, "LNW")) algo = DPX_ALGO_LNW; else if (!strcmp(a, "ANW")) algo = DPX_ALGO_ANW;
            else if (!strcmp(a, "LSW")) algo = DPX_ALGO_LSW; else if (!strcmp(a, "BSW")) algo = DPX_ALGO_BSW;
            else if (!strcmp(a, "ABSW")) algo = DPX_ALGO_ABSW;        // affine banded SW (-open, -extend, -band)
            else { fprintf(stderr, "unknown -algo %s\n", a); exit(EXIT_FAILURE); }
        }
    }
    if (!pairFileName && fastx) pairFileName = fastx;
    if (!pairFileName) { fprintf(stderr, "missing -pairs <InSeqFile>\n"); exit(EXIT_FAILURE); }

    printf("Parsing input file: %s\n", pairFileName);
    const auto t0 = std::chrono::steady_clock::now();
    printf("Pair # | Score\n");

    // The file goes to the GPU as is: parser (newline scan, seqPair index, 2-bit pack), alignment, backtrack and the formatting of
    // the stdout blocks all run on the device; one text buffer comes back (reference: parseInput + the per-pair loop, main.cpp:157-252).
    dpx_params p = dpxhost::make_params(algo, matchWeight, mismatchWeight, gapOpenWeight, gapExtendWeight, band);
    if (scores_only) p.flags = DPX_OUT_SCORE | DPX_OUT_END_COORDS;
    char* text = nullptr; size_t text_bytes = 0;
    int st;
    if (!devices.empty() && !long_pairs && !all_maxima) {
        // several GPUs, one process: parse on the host (the parser also leaves the packed 2-bit copy every shard uploads from),
        // then one multi-device batch call; the text blocks come back in pair order
        dpx_multi* m = nullptr;
        st = dpx_create_multi(&m, devices.data(), (int)devices.size());
        if (st != DPX_OK) { fprintf(stderr, "dpxalign: cannot create the device contexts: %s\n", dpx_strerror(st)); exit(1); }
        dpx_seq_pair* idx = nullptr; char* seqs = nullptr; dpx_input_info info{};
        st = fastx ? dpx_parse_fastx(fastx, fastx2, &idx, &seqs, &info) : dpx_parse_input(pairFileName, &idx, &seqs, &info);
        if (st == DPX_OK) st = dpx_multi_align_batch_text(m, &p, seqs, info.numBytes, idx, info.numPairs, 0, nullptr, nullptr, &text, &text_bytes);
        if (st == DPX_ERR_IO) { fprintf(stderr, "Could not open file: %s\n", pairFileName); exit(1); }
        if (st == DPX_ERR_FORMAT) { fprintf(stderr, "Number of lines not a multiple of 3: %s\n", pairFileName); exit(1); }
        if (st != DPX_OK) { fprintf(stderr, "dpxalign: %s (%s)\n", dpx_strerror(st), dpx_multi_last_error(m)); exit(1); }
        dpx_free(idx); dpx_free(seqs);
        fwrite(text, 1, text_bytes, stdout); dpx_free(text);
        dpx_destroy_multi(m);
        const long long usec = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        printf("Elapsed time (usec): %lld\n", usec);
        printf("Cleaning up\n");
        return 0;
    }
    dpx_ctx* ctx = dpxhost::engine();
    if (fastx || long_pairs || all_maxima) {
        // host-side parsers: the project's 3-line records (parseInput) or FASTA / FASTQ records, then one batch — or, with -long, one
        // checkpointed alignment per pair, printed as LinearSmithWaterman::print_results does (c++/LinearSmithWaterman.cpp:240-288)
        dpx_seq_pair* idx = nullptr; char* seqs = nullptr; dpx_input_info info{};
        st = fastx ? dpx_parse_fastx(fastx, fastx2, &idx, &seqs, &info) : dpx_parse_input(pairFileName, &idx, &seqs, &info);
        if (st == DPX_OK && long_pairs) {
            if (algo != DPX_ALGO_LSW) { fprintf(stderr, "-long aligns with LinearSmithWaterman only\n"); exit(EXIT_FAILURE); }
            for (size_t k = 0; k < info.numPairs && st == DPX_OK; ++k) {
                int32_t score = 0; char* lines = nullptr; size_t len = 0;
                st = dpx_align_long_pair_strings(ctx, &p, seqs + idx[k].referenceIdx, (size_t)idx[k].referenceSize, seqs + idx[k].queryIdx,
                                                 (size_t)idx[k].querySize, &score, nullptr, nullptr, nullptr, nullptr, &lines, &len, nullptr);
                if (st != DPX_OK) break;
                printf("%zu | %d\n", k, score);
                if (score == 0 || scores_only) { if (!scores_only) fputs("\n\n\n", stdout); }
                else for (int l = 0; l < 3; ++l) { fwrite(lines + (size_t)l * (len + 1), 1, len, stdout); fputc('\n', stdout); }
                dpx_free(lines);
            }
        } else if (st == DPX_OK && all_maxima) {
            if (algo != DPX_ALGO_LSW) { fprintf(stderr, "-all is a LinearSmithWaterman mode\n"); exit(EXIT_FAILURE); }
            st = dpx_align_batch_text_all(ctx, &p, seqs, info.numBytes, idx, info.numPairs, 0, &text, &text_bytes, nullptr);
        } else if (st == DPX_OK) {
            st = dpx_align_batch_text(ctx, &p, seqs, info.numBytes, idx, info.numPairs, 0, nullptr, nullptr, &text, &text_bytes);
        }
        dpx_free(idx); dpx_free(seqs);
    } else {
        st = dpx_align_file_text(ctx, &p, pairFileName, 0, &text, &text_bytes, nullptr);
    }
    if (st == DPX_ERR_IO) { fprintf(stderr, "Could not open file: %s\n", pairFileName); exit(1); }                      // parseInput.cpp:12-15
    if (st == DPX_ERR_FORMAT) { fprintf(stderr, "Number of lines not a multiple of 3: %s\n", pairFileName); exit(1); }  // :38-41
    if (st != DPX_OK) { fprintf(stderr, "dpxalign: %s (%s)\n", dpx_strerror(st), dpx_last_error(ctx)); exit(1); }
    if (text) { fwrite(text, 1, text_bytes, stdout); dpx_free(text); }

    const long long usec = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    printf("Elapsed time (usec): %lld\n", usec);
    printf("Cleaning up\n");
    return 0;
}
