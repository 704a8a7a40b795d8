// The reference's own per-pair loop (c++/main.cpp:237-252), compiled unchanged against the shim classes:
// construct the aligner on (&sequences[referenceIdx], &sequences[queryIdx], i, weights) and call align().
#include <cstring>
#include "parseInput.h"
#include "LinearSmithWaterman.h"
#include "LinearNeedlemanWunsch.h"
#include "AffineNeedlemanWunsch.h"
#include "BandedSmithWaterman.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: example_per_pair LNW|LSW|ANW|BSW <file> [limit]\n"); return 2; }
    seqPair* idx; char* sequences;
    inputInfo info = parseInput(argv[2], idx, sequences);
    size_t n = info.numPairs;
    if (argc > 3 && (size_t)atol(argv[3]) < n) n = (size_t)atol(argv[3]);
    for (size_t i = 0; i < n; ++i) {
        const char* r = &sequences[idx[i].referenceIdx];
        const char* q = &sequences[idx[i].queryIdx];
        if (!strcmp(argv[1], "LNW")) { LinearNeedlemanWunsch a(r, q, (int)i, 3, -1, -2); a.align(); }
        else if (!strcmp(argv[1], "LSW")) { LinearSmithWaterman a(r, q, (int)i, 3, -1, -2); a.align(); }
        else if (!strcmp(argv[1], "ANW")) { AffineNeedlemanWunsch a(r, q, (int)i, 3, -1, -3, -1); a.align(); }
        else { BandedSmithWaterman a(r, q, 3, -1, -2, (int)i, 1 << 20); a.align(); }
    }
    cleanupParsedFile(idx, sequences);
    return 0;
}
