// parseInput.h — same types and signatures as the reference parser (c++/parseInput.h:9-35).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

struct inputInfo {
    size_t numPairs;
    size_t numBytes;
    size_t numCells;
    size_t minReferenceLength;
    size_t minQueryLength;
    size_t maxReferenceLength;
    size_t maxQueryLength;
    double avgReferenceLength;
    double avgQueryLength;
};

struct seqPair {
    int referenceIdx;
    int referenceSize;
    int queryIdx;
    int querySize;
};

inputInfo parseInput(const char* pairFileName, seqPair*& sequence_indices, char*& sequences);
void printParsedFile(const size_t numPairs, const seqPair* sequence_indices, const char* sequences);
void cleanupParsedFile(seqPair* sequence_indices, char* sequences);
