#include "GpuAligner.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <vector>

namespace dpxhost {

dpx_ctx* engine() {
    static dpx_ctx* ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (const char* e = getenv("DPX_DEVICE")) dev = atoi(e);
        int st = dpx_create(&ctx, dev);
        if (st != DPX_OK) {
            fprintf(stderr, "dpxalign: cannot create a device context: %s\n", dpx_strerror(st));
            exit(1);
        }
    });
    return ctx;
}

static std::mutex g_engine_lock;      // one ctx: calls are serialised (aligner objects stay thread-confined)

void GpuAligner::run() {
    if (done) return;
    const size_t R = reference_str.size(), Q = query_str.size();
    std::vector<char> blob(R + Q + 2);
    if (R) memcpy(blob.data(), reference_str.data(), R);
    blob[R] = '\0';
    if (Q) memcpy(blob.data() + R + 1, query_str.data(), Q);
    blob[R + 1 + Q] = '\0';
    dpx_seq_pair pr{0, (int32_t)R, (int32_t)(R + 1), (int32_t)Q};
    int32_t rc[2] = {0, 0};
    char* strings = nullptr; size_t* offs = nullptr;
    int st;
    {
        std::lock_guard<std::mutex> g(g_engine_lock);
        st = dpx_align_batch(engine(), &params, blob.data(), blob.size(), &pr, 1, &score, rc, &strings, &offs);
        if (st != DPX_OK) fprintf(stderr, "dpxalign: %s (%s)\n", dpx_strerror(st), dpx_last_error(engine()));
    }
    if (st != DPX_OK) exit(1);
    end_row = rc[0]; end_col = rc[1];
    reference_sequence = strings + offs[0];
    pair_relation = strings + offs[1];
    query_sequence = strings + offs[2];
    dpx_free(strings); dpx_free(offs);
    done = true;
}

void GpuAligner::print_matrix() {
    std::cout << "Reference: " << reference_str << " Size: " << reference_str.size() << "\n";
    std::cout << "Query: " << query_str << " Size: " << query_str.size() << "\n";
    std::cout << "[score matrix is not materialised by the GPU engine]" << std::endl;
}

void GpuAligner::print_results() {
    run();
    // exactly the reference's bytes: printf("%d | ", pairNum); cout << score << "\n"; then the three lines
    // (reference c++/LinearNeedlemanWunsch.cpp:207-213, c++/LinearSmithWaterman.cpp:252-279).
    std::string out = std::to_string(pairNum) + " | " + std::to_string(score) + "\n" + reference_sequence + "\n" +
                      pair_relation + "\n" + query_sequence + "\n";
    fwrite(out.data(), 1, out.size(), stdout);
    fflush(stdout);
}

}  // namespace dpxhost
