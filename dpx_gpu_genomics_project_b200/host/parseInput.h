// parseInput.h — the reference parser's names on top of the C ABI's types: `seqPair` and `inputInfo` ARE dpx_seq_pair and
// dpx_input_info (include/dpxalign.h, which keeps the reference's field names and layout, c++/parseInput.h:9-29), so arrays
// pass between the shims and the library without conversion; the three entry points keep the reference's signatures (:31-35).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "../../include/dpxalign.h"

typedef dpx_input_info inputInfo;      // numPairs, numBytes, numCells, min/max/avg Reference/Query Length
typedef dpx_seq_pair seqPair;          // referenceIdx, referenceSize, queryIdx, querySize

inputInfo parseInput(const char* pairFileName, seqPair*& sequence_indices, char*& sequences);
void printParsedFile(const size_t numPairs, const seqPair* sequence_indices, const char* sequences);
void cleanupParsedFile(seqPair* sequence_indices, char* sequences);
