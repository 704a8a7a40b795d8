// TEST INFRASTRUCTURE: drives the host traceback utilities (dpx_gpu_genomics_project_b200/host/backtrack.{h,cpp}) the way the
// reference's CUDA mains drive c++/backtrack.cpp: it fills plain direction matrices on the CPU with the reference's recurrences
// (c++/LinearNeedlemanWunsch.cpp:97-132, c++/AffineNeedlemanWunsch.cpp:175-238, c++/LinearSmithWaterman.cpp:75-110 + :145-157)
// and lets backtrackMultiNW / backtrackANW / backtrackSW print.  The output must equal the committed golden text of the reference
// classes (tests/golden/*.out.txt).    usage: backtrack_driver LNW|ANW|LSW file match mismatch open extend
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "backtrack.h"

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const std::string algo = argv[1];
    const int ma = atoi(argv[3]), mi = atoi(argv[4]), go = atoi(argv[5]), ge = atoi(argv[6]);
    std::ifstream in(argv[2]);
    std::string header, ref, qry;
    for (int pairNum = 0; std::getline(in, header) && std::getline(in, ref) && std::getline(in, qry); ++pairNum) {
        const int R = (int)ref.size(), Q = (int)qry.size();
        const size_t W = (size_t)R + 1, N = W * ((size_t)Q + 1);
        std::vector<int> H(N, 0);
        std::vector<directionMain> D(N, NONE_MAIN);
        if (algo == "LNW") {
            for (int i = 1; i <= Q; ++i) { H[i * W] = i * go; D[i * W] = QUERY_DELETION; }
            for (int j = 1; j <= R; ++j) { H[j] = j * go; D[j] = QUERY_INSERTION; }
            for (int i = 1; i <= Q; ++i)
                for (int j = 1; j <= R; ++j) {
                    const bool eq = qry[i - 1] == ref[j - 1];
                    int m = H[(i - 1) * W + j - 1] + (eq ? ma : mi); directionMain d = eq ? MATCH : MISMATCH;
                    const int up = H[(i - 1) * W + j] + go, left = H[i * W + j - 1] + go;
                    if (up >= m) { m = up; d = QUERY_DELETION; }
                    if (left >= m) { m = left; d = QUERY_INSERTION; }
                    H[i * W + j] = m; D[i * W + j] = d;
                }
            backtrackMultiNW(D.data(), ref.c_str(), R, qry.c_str(), Q, pairNum, H[(size_t)Q * W + R]);
        } else if (algo == "ANW") {
            std::vector<int> Dm(N, 0), Im(N, 0);
            std::vector<directionIndel> dI(N, NONE_INDEL), dD(N, NONE_INDEL);
            for (int i = 1; i <= Q; ++i) { H[i * W] = go + i * ge; D[i * W] = QUERY_DELETION; }
            for (int j = 1; j <= R; ++j) { H[j] = go + j * ge; D[j] = QUERY_INSERTION; }
            for (int i = 1; i <= Q; ++i)
                for (int j = 1; j <= R; ++j) {
                    const size_t c = (size_t)i * W + j;
                    const int dOpen = H[c - W] + go + ge, dExt = Dm[c - W] + ge;
                    if (i == 1 || dOpen >= dExt) { Dm[c] = dOpen; dD[c] = GAP_OPEN; } else { Dm[c] = dExt; dD[c] = GAP_EXTEND; }
                    const int iOpen = H[c - 1] + go + ge, iExt = Im[c - 1] + ge;
                    if (j == 1 || iOpen >= iExt) { Im[c] = iOpen; dI[c] = GAP_OPEN; } else { Im[c] = iExt; dI[c] = GAP_EXTEND; }
                    const bool eq = qry[i - 1] == ref[j - 1];
                    int m = H[c - W - 1] + (eq ? ma : mi); directionMain d = eq ? MATCH : MISMATCH;
                    if (Dm[c] >= m) { m = Dm[c]; d = QUERY_DELETION; }
                    if (Im[c] >= m) { m = Im[c]; d = QUERY_INSERTION; }
                    H[c] = m; D[c] = d;
                }
            printf("%d | %d\n", pairNum, H[(size_t)Q * W + R]);
            backtrackANW(D.data(), dI.data(), dD.data(), ref.c_str(), R, qry.c_str(), Q);
        } else {
            int best = 0, bi = 0, bj = 0;
            for (int i = 1; i <= Q; ++i)
                for (int j = 1; j <= R; ++j) {
                    const size_t c = (size_t)i * W + j;
                    const bool eq = qry[i - 1] == ref[j - 1];
                    const int up = H[c - W] + go, left = H[c - 1] + go, diag = H[c - W - 1] + (eq ? ma : mi);
                    const int t = std::max(up, std::max(left, diag)), h = std::max(0, t);
                    H[c] = h;
                    if (h > 0) D[c] = (up == h) ? QUERY_DELETION : (left == h) ? QUERY_INSERTION : (eq ? MATCH : MISMATCH);   // NONE_MAIN <=> H == 0
                }
            for (int i = 0; i <= Q; ++i) for (int j = 0; j <= R; ++j) if (H[(size_t)i * W + j] > best) { best = H[(size_t)i * W + j]; bi = i; bj = j; }
            printf("%d | %d\n", pairNum, best);
            if (best == 0) fputs("\n\n\n", stdout);
            else backtrackSW(bi, bj, (int)W, D.data(), ref.c_str(), qry.c_str());
        }
    }
    return 0;
}
