// backtrack.cpp — see backtrack.h.  Host utilities for callers with their own direction matrices; not on the library's path.
// One shared walker emits the three lines right to left into buffers of queryLength + referenceLength characters (the reference
// prepends to std::string, O(L^2)); an unknown direction ends the process with exit(1), as the reference does (:71,134,199,277).
#include "backtrack.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

namespace {

std::mutex g_print;     // the reference serialises multi-threaded printing (c++/printLock.cpp:3-11)

struct Lines {
    std::string ref, rel, qry; size_t pos;
    explicit Lines(size_t cap) : ref(cap, ' '), rel(cap, ' '), qry(cap, ' '), pos(cap) {}
    void push(char r, char l, char q) { --pos; ref[pos] = r; rel[pos] = l; qry[pos] = q; }
    void print() const { printf("%s\n%s\n%s\n", ref.c_str() + pos, rel.c_str() + pos, qry.c_str() + pos); }
};

// one move of the linear walks; false on an unknown direction
bool step(directionMain d, const char* r, const char* q, int& i, int& j, Lines& out) {
    switch (d) {
        case MATCH:           out.push(r[j - 1], '*', q[i - 1]); --i; --j; return true;
        case MISMATCH:        out.push(r[j - 1], '|', q[i - 1]); --i; --j; return true;
        case QUERY_DELETION:  out.push('_', ' ', q[i - 1]); --i; return true;
        case QUERY_INSERTION: out.push(r[j - 1], ' ', '_'); --j; return true;
        default: return false;
    }
}

}  // namespace

void printMatrix(const int* memo, const int cols, const int rows) {
    for (int row = 0; row < rows; ++row) { for (int col = 0; col < cols; ++col) printf(" %4d ", memo[(size_t)row * cols + col]); printf("\n"); }
}

void printBacktrackMatrix(const directionMain* memo, const int cols, const int rows) {
    for (int row = 0; row < rows; ++row) { for (int col = 0; col < cols; ++col) printf(" %4d ", (int)memo[(size_t)row * cols + col]); printf("\n"); }
}

void backtrackNW(const directionMain* memo, const char* r, const int R, const char* q, const int Q) {
    const size_t W = (size_t)R + 1;
    int i = Q, j = R;
    Lines out((size_t)Q + R);
    while (i != 0 || j != 0)
        if (!step(memo[(size_t)i * W + j], r, q, i, j, out)) exit(1);
    out.print();
}

void backtrackMultiNW(const directionMain* memo, const char* r, const int R, const char* q, const int Q, const int pairNum, const int score) {
    const size_t W = (size_t)R + 1;
    int i = Q, j = R;
    Lines out((size_t)Q + R);
    while (i != 0 || j != 0)
        if (!step(memo[(size_t)i * W + j], r, q, i, j, out)) {
            { std::lock_guard<std::mutex> g(g_print); printf("Exiting(1) backtrack: %d\n", pairNum); }
            exit(1);
        }
    std::lock_guard<std::mutex> g(g_print);
    printf("%d | %d\n", pairNum, score);
    out.print();
}

void backtrackSW(int i, int j, const int numCols, const directionMain* memo, const char* r, const char* q) {
    Lines out((size_t)i + (size_t)j);
    while (i > 0 && j > 0 && memo[(size_t)i * numCols + j] != NONE_MAIN)
        if (!step(memo[(size_t)i * numCols + j], r, q, i, j, out)) exit(1);
    out.print();
}

void backtrackANW(const directionMain* scoring, const directionIndel* ins, const directionIndel* del, const char* r, const int R, const char* q, const int Q) {
    const size_t W = (size_t)R + 1;
    int i = Q, j = R;
    currentMatrixPosition at = SCORING;
    Lines out((size_t)Q + R);
    while (i != 0 && j != 0) {
        const size_t c = (size_t)i * W + j;
        if (at == SCORING) {
            switch (scoring[c]) {
                case MATCH: case MISMATCH: step(scoring[c], r, q, i, j, out); break;
                case QUERY_DELETION: at = DELETION; break;
                case QUERY_INSERTION: at = INSERTION; break;
                default: exit(1);
            }
        } else if (at == INSERTION) {                     // the direction is read BEFORE the move (:309-332)
            if (ins[c] == GAP_OPEN) at = SCORING; else if (ins[c] != GAP_EXTEND) exit(1);
            out.push(r[j - 1], ' ', '_'); --j;
        } else {
            if (del[c] == GAP_OPEN) at = SCORING; else if (del[c] != GAP_EXTEND) exit(1);
            out.push('_', ' ', q[i - 1]); --i;
        }
    }
    while (i > 0) { out.push('_', ' ', q[i - 1]); --i; }  // :366-371
    while (j > 0) { out.push(r[j - 1], ' ', '_'); --j; }  // :373-378
    out.print();
}
