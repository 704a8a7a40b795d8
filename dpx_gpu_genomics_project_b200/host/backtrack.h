// backtrack.h — the reference's host traceback interface (c++/backtrack.h:14-53), kept for source compatibility of callers that
// bring their OWN direction matrices (the reference's CUDA mains do: cuda/LNW/LinearNeedlemanWunsch.cu:322, cuda/LinearSmithWaterman.cu:316,
// cuda/AffineNeedlemanWunsch.cu:357-391).  libdpxalign never calls these: its traceback runs on the GPU over packed 2/4-bit codes
// (csrc/backtrack.cuh, pairwf.cuh, band.cuh) and returns the strings through the C ABI.  Same signatures, same stdout bytes;
// matrices are row-major (queryLength + 1) x (referenceLength + 1), as the reference's.
#pragma once

enum directionMain { NONE_MAIN, MATCH, MISMATCH, QUERY_INSERTION, QUERY_DELETION };
enum directionIndel { NONE_INDEL, GAP_OPEN, GAP_EXTEND };
enum currentMatrixPosition { SCORING, INSERTION, DELETION };

void printMatrix(const int* memo, const int referenceLength, const int queryLength);
void printBacktrackMatrix(const directionMain* memo, const int referenceLength, const int queryLength);

// c++/backtrack.cpp:21-81: from (queryLength, referenceLength) until both indices are 0; prints REF / REL / QRY.
void backtrackNW(const directionMain* backtrackMemo, const char* referenceString, const int referenceLength, const char* queryString, const int queryLength);
// :146-212: the same walk, printed under the print lock after a "<pairNum> | <score>" line.
void backtrackMultiNW(const directionMain* backtrackMemo, const char* referenceString, const int referenceLength, const char* queryString, const int queryLength,
                      const int pairNum, const int score);
// :83-144: from the given cell while both indices are > 0 and the cell is not NONE_MAIN.
void backtrackSW(int currentMemoRow, int currentMemoCol, const int numCols, const directionMain* backtrackMemo, const char* referenceString, const char* queryString);
// :214-356: Gotoh's three-state walk, then the remaining rows as deletions and the remaining columns as insertions.
void backtrackANW(const directionMain* scoringBacktrack, const directionIndel* queryInsertionBacktrack, const directionIndel* queryDeletionBacktrack,
                  const char* referenceString, const int referenceLength, const char* queryString, const int queryLength);
