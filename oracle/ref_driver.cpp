// TEST INFRASTRUCTURE ONLY (oracle/): driver that links the UNMODIFIED reference
// aligner classes, compiled in place from /root/reference/c++ (never copied), so the
// real reference arithmetic can (a) pin the C restatement in dpx_oracle.c, (b) generate
// the golden fixtures in tests/golden/, (c) serve as the "reference" CPU baseline in
// bench.py.  Nothing in the product path (libdpxalign.so, the host shims) links or
// executes this file.
//
// Why our own driver instead of the reference's c++/main.cpp:
//   * main.cpp:169 computes numBatches with integer division, so the threaded path
//     aligns nothing for < 400 pairs and drops tail pairs;
//   * main.cpp:249 (ANW sequential branch) has a syntax error (`numPairs/;`);
//   * the algorithm is a compile-time #define (main.cpp:22-24).
// The per-pair call sequence is exactly main.cpp:237-252: construct the aligner on
// (&sequences[referenceIdx], &sequences[queryIdx], i, weights) and call align().
// Linear aligners receive the -open weight as their gap weight (main.cpp:238,244).
//
// usage: ref_align -algo LNW|ANW|LSW -pairs F [-match m] [-mismatch x] [-open g]
//                  [-extend e] [-threads T] [-limit N] [-noheader]
// stdout: "Parsing input file: ..", "Pair # | Score", the blocks, "Elapsed..", "Cleaning up"
//         (exactly main.cpp:153,165,257,260) unless -noheader.
// stderr: "ALIGN_USEC <n> CELLS <n> PAIRS <n> THREADS <n>" (align loop only, parse excluded).
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <pthread.h>
#include <sys/time.h>

#include "parseInput.h"
#include "LinearSmithWaterman.h"
#include "LinearNeedlemanWunsch.h"
#include "AffineNeedlemanWunsch.h"

namespace {

enum Algo { A_LNW, A_ANW, A_LSW };

struct Job {
    Algo algo;
    const char* sequences;
    const seqPair* idx;
    size_t begin, end;
    int match, mismatch, open, extend;
};

void* run_range(void* p) {
    const Job* j = static_cast<const Job*>(p);
    for (size_t i = j->begin; i < j->end; ++i) {
        const char* r = &j->sequences[j->idx[i].referenceIdx];
        const char* q = &j->sequences[j->idx[i].queryIdx];
        switch (j->algo) {
            case A_LNW: { LinearNeedlemanWunsch a(r, q, (int)i, j->match, j->mismatch, j->open); a.align(); break; }
            case A_LSW: { LinearSmithWaterman   a(r, q, (int)i, j->match, j->mismatch, j->open); a.align(); break; }
            case A_ANW: { AffineNeedlemanWunsch a(r, q, (int)i, j->match, j->mismatch, j->open, j->extend); a.align(); break; }
        }
    }
    return nullptr;
}

uint64_t now_usec() {
    timeval tv; gettimeofday(&tv, nullptr);
    return tv.tv_sec * (uint64_t)1000000 + tv.tv_usec;
}

}  // namespace

int main(int argc, char** argv) {
    const char* file = nullptr; Algo algo = A_LNW;
    int match = 3, mismatch = -1, open = -4, extend = -1;   // defaults of main.cpp:128-132
    int threads = 1; long limit = -1; bool header = true;
    for (int i = 1; i < argc; ++i) {
        auto next = [&](const char* f) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", f); exit(2); }
            return argv[++i];
        };
        if      (!strcmp(argv[i], "-pairs"))    file = next("-pairs");
        else if (!strcmp(argv[i], "-algo"))   { const char* a = next("-algo");
                                                if (!strcmp(a, "LNW")) algo = A_LNW; else if (!strcmp(a, "ANW")) algo = A_ANW;
                                                else if (!strcmp(a, "LSW")) algo = A_LSW; else { fprintf(stderr, "bad -algo %s\n", a); exit(2); } }
        else if (!strcmp(argv[i], "-match"))    match = atoi(next("-match"));
        else if (!strcmp(argv[i], "-mismatch")) mismatch = atoi(next("-mismatch"));
        else if (!strcmp(argv[i], "-open"))     open = atoi(next("-open"));
        else if (!strcmp(argv[i], "-extend"))   extend = atoi(next("-extend"));
        else if (!strcmp(argv[i], "-threads"))  threads = atoi(next("-threads"));
        else if (!strcmp(argv[i], "-limit"))    limit = atol(next("-limit"));
        else if (!strcmp(argv[i], "-noheader")) header = false;
        else { fprintf(stderr, "unknown flag %s\n", argv[i]); exit(2); }
    }
    if (!file) { fprintf(stderr, "usage: ref_align -algo LNW|ANW|LSW -pairs F ...\n"); return 2; }
    if (threads < 1) threads = 1;

    if (header) printf("Parsing input file: %s\n", file);
    seqPair* idx; char* sequences;
    inputInfo info = parseInput(file, idx, sequences);
    size_t n = info.numPairs;
    if (limit >= 0 && (size_t)limit < n) n = (size_t)limit;

    size_t cells = 0;
    for (size_t i = 0; i < n; ++i) cells += (size_t)idx[i].referenceSize * (size_t)idx[i].querySize;

    if (header) printf("Pair # | Score\n");
    fflush(stdout);
    uint64_t t0 = now_usec();
    if (threads == 1) {
        Job j{algo, sequences, idx, 0, n, match, mismatch, open, extend};
        run_range(&j);
    } else {
        std::vector<Job> jobs(threads);
        std::vector<pthread_t> tids(threads);
        for (int t = 0; t < threads; ++t) {
            jobs[t] = Job{algo, sequences, idx, n * t / threads, n * (t + 1) / threads, match, mismatch, open, extend};
            pthread_create(&tids[t], nullptr, run_range, &jobs[t]);
        }
        for (int t = 0; t < threads; ++t) pthread_join(tids[t], nullptr);
    }
    fflush(stdout);
    uint64_t dt = now_usec() - t0;
    if (header) { printf("Elapsed time (usec): %lld\n", (long long)dt); printf("Cleaning up\n"); }
    fprintf(stderr, "ALIGN_USEC %llu CELLS %zu PAIRS %zu THREADS %d\n", (unsigned long long)dt, cells, n, threads);
    cleanupParsedFile(idx, sequences);
    return 0;
}
