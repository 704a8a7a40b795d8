// parseInput.cpp — the reference's parser entry points (c++/parseInput.cpp:9-143) on top of dpx_parse_input.
// Error behaviour of the reference is kept: a message on stderr and exit(1) (:14, :40).
#include "parseInput.h"

static_assert(sizeof(int) == sizeof(int32_t), "the reference's seqPair holds four ints (c++/parseInput.h:22-29)");

inputInfo parseInput(const char* pairFileName, seqPair*& sequence_indices, char*& sequences) {
    inputInfo in;
    sequence_indices = nullptr; sequences = nullptr;
    const int st = dpx_parse_input(pairFileName, &sequence_indices, &sequences, &in);
    if (st == DPX_ERR_IO) { fprintf(stderr, "Could not open file: %s\n", pairFileName); exit(1); }
    if (st == DPX_ERR_FORMAT) { fprintf(stderr, "Number of lines not a multiple of 3: %s\n", pairFileName); exit(1); }
    if (st != DPX_OK) { fprintf(stderr, "parseInput: %s\n", dpx_strerror(st)); exit(1); }
    return in;
}

void printParsedFile(const size_t numPairs, const seqPair* idx, const char* sequences) {
    for (size_t i = 0; i < numPairs; ++i)
        printf("Pair: %zu Reference: %c, Reference Size: %d, Query: %c, Query Size: %d\n", i,
               sequences[idx[i].referenceIdx], idx[i].referenceSize, sequences[idx[i].queryIdx], idx[i].querySize);
}

void cleanupParsedFile(seqPair* sequence_indices, char* sequences) {
    dpx_free(sequence_indices);
    dpx_free(sequences);
}
