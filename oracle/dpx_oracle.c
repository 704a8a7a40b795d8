/* dpx_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's pairwise-alignment hot path
 * (mickgordinier/DPX_GPU_Genomics_Project, c++/).  It exists to CHECK the CUDA product
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).  Nothing in the
 * product path (libdpxalign.so, host shims) may link, import or execute it.
 *
 * Parity status (see tests/test_oracle_vs_reference.py and tests/golden/):
 *   LNW, ANW, LSW : PINNED — byte-identical text to the compiled, unmodified reference
 *                   classes (oracle/_ref/ref_align) on the committed golden fixtures and,
 *                   when /root/reference is present, on fresh random/adversarial inputs.
 *   BSW (banded)  : PARITY UNPINNED against C++ — c++/BandedSmithWaterman.cpp is not
 *                   executable (uninitialised band_width .h:16,51-57; size_t underflow
 *                   .cpp:79-83; shadowed loop .cpp:87,92).  Repaired semantics:
 *                   LinearSmithWaterman restricted to |i-j| <= W, everything else 0.  Pins:
 *                   W >= max(Q,R) degenerates to the pinned LSW; scores equal the Python
 *                   prototype python/LinearBandedSmithWaterman.py:71 run with BAND = W+1
 *                   (fixtures in tests/golden/bsw_python_scores.json).
 *
 *   ABSW (affine banded SW): NOT A REFERENCE ALGORITHM — the reference only names it as a TODO
 *                   (python/LinearBandedSmithWaterman.py:8 "BSW evidently usually uses affine gap
 *                   penalties").  Definition adopted: local Gotoh restricted to
 *                   |i-j| <= W, built from the reference's own pieces — D/I recurrences and GAP_OPEN tie rule
 *                   of AffineNeedlemanWunsch (c++/AffineNeedlemanWunsch.cpp:185-213), ReLU / UP > LEFT > DIAG /
 *                   first-strict-max end cell / stop-at-zero walk of LinearSmithWaterman.  Pins: gap_open = 0
 *                   degenerates byte-for-byte to the (pinned) linear banded SW with gap = gap_extend; scores
 *                   equal an independent numpy three-state local alignment (tests/test_oracle.py).
 *
 * Each function cites the reference file:line it follows.  Matrices are (Q+1)x(R+1)
 * row-major, rows = query, cols = reference, all arithmetic int32 (as the reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define ORC_LNW 0
#define ORC_ANW 1
#define ORC_LSW 2
#define ORC_BSW 3
#define ORC_ABSW 4   /* affine banded Smith-Waterman: NOT in the reference (see absw_pair) */

typedef struct { int32_t referenceIdx, referenceSize, queryIdx, querySize; } orc_pair; /* c++/parseInput.h:22-29 */
typedef struct { int32_t algo, match, mismatch, gap_open, gap_extend, band; } orc_params;

/* direction codes (ours; the reference uses enums, c++/backtrack.h:14-27) */
enum { D_NONE = 0, D_DIAG = 1, D_UP = 2 /* QUERY_DELETION */, D_LEFT = 3 /* QUERY_INSERTION */ };

/* FakeDPX::__vibmax_s32, c++/FakeDPX.cpp:145-153: returns max, pred = (a >= b). */
static inline int32_t vibmax_s32(int32_t a, int32_t b, int* pred) {
    if (a >= b) { *pred = 1; return a; }
    *pred = 0; return b;
}

/* Right-to-left string builder: the reference prepends one char per step
 * (e.g. c++/LinearNeedlemanWunsch.cpp:157-159); we fill from the end of a buffer. */
typedef struct { char *ref, *rel, *qry; int64_t cap, pos; } sb3;
static void sb_init(sb3* s, char* r, char* l, char* q, int64_t cap) { s->ref = r; s->rel = l; s->qry = q; s->cap = cap; s->pos = cap; }
static inline void sb_push(sb3* s, char r, char l, char q) { s->pos--; s->ref[s->pos] = r; s->rel[s->pos] = l; s->qry[s->pos] = q; }
static int64_t sb_finish(sb3* s) {
    int64_t len = s->cap - s->pos;
    memmove(s->ref, s->ref + s->pos, (size_t)len); s->ref[len] = 0;
    memmove(s->rel, s->rel + s->pos, (size_t)len); s->rel[len] = 0;
    memmove(s->qry, s->qry + s->pos, (size_t)len); s->qry[len] = 0;
    return len;
}

/* ---------------------------------------------------------------------------------
 * LinearNeedlemanWunsch: init_matrix c++/LinearNeedlemanWunsch.cpp:9-42,
 * score_matrix :89-135, backtrack :137-223.  Tie precedence LEFT > UP > DIAG. */
static int64_t lnw_pair(const orc_params* p, const char* r, int R, const char* q, int Q,
                        int32_t* score, char* o_ref, char* o_rel, char* o_qry) {
    size_t W = (size_t)R + 1;
    int32_t* H = (int32_t*)malloc(sizeof(int32_t) * W * ((size_t)Q + 1));
    uint8_t* D = (uint8_t*)malloc(W * ((size_t)Q + 1));
    const int32_t g = p->gap_open;            /* linear aligners get -open as gap: c++/main.cpp:238,244 */
    H[0] = 0; D[0] = D_NONE;
    for (int i = 1; i <= Q; ++i) { H[i * W] = i * g; D[i * W] = D_UP; }      /* :31-34 */
    for (int j = 1; j <= R; ++j) { H[j] = j * g; D[j] = D_LEFT; }            /* :38-41 */
    for (int i = 1; i <= Q; ++i) {
        for (int j = 1; j <= R; ++j) {
            int pred;
            int32_t diag = H[(i - 1) * W + (j - 1)] + (q[i - 1] == r[j - 1] ? p->match : p->mismatch); /* :108-116 */
            uint8_t d = D_DIAG;
            int32_t up = H[(i - 1) * W + j] + g;                              /* :119 */
            int32_t left = H[i * W + (j - 1)] + g;                            /* :120 */
            int32_t m = vibmax_s32(up, diag, &pred);  if (pred) d = D_UP;     /* :122-123 */
            m = vibmax_s32(left, m, &pred);           if (pred) d = D_LEFT;   /* :125-126 */
            H[i * W + j] = m; D[i * W + j] = d;
        }
    }
    *score = H[(size_t)Q * W + R];                                            /* :209 */
    int64_t len = -1;
    if (o_ref) {
        sb3 s; sb_init(&s, o_ref, o_rel, o_qry, (int64_t)Q + R);
        int i = Q, j = R;
        while (i != 0 || j != 0) {                                            /* :151 */
            switch (D[i * W + j]) {
                case D_DIAG: sb_push(&s, r[j - 1], q[i - 1] == r[j - 1] ? '*' : '|', q[i - 1]); --i; --j; break; /* :156-173 */
                case D_UP:   sb_push(&s, '_', ' ', q[i - 1]); --i; break;     /* :176-181 */
                case D_LEFT: sb_push(&s, r[j - 1], ' ', '_'); --j; break;     /* :184-189 */
                default: abort();
            }
        }
        len = sb_finish(&s);
    }
    free(H); free(D);
    return len;
}

/* ---------------------------------------------------------------------------------
 * AffineNeedlemanWunsch (Gotoh): init_matrix c++/AffineNeedlemanWunsch.cpp:12-54,
 * score_matrix :167-240, backtrack :242-403.  A k-long gap costs open + k*extend.
 * dir byte: bits0-1 = dirH (D_DIAG/D_UP/D_LEFT), bit2 = dirD is GAP_OPEN, bit3 = dirI is GAP_OPEN. */
static int64_t anw_pair(const orc_params* p, const char* r, int R, const char* q, int Q,
                        int32_t* score, char* o_ref, char* o_rel, char* o_qry) {
    size_t W = (size_t)R + 1, N = W * ((size_t)Q + 1);
    int32_t* H  = (int32_t*)calloc(N, sizeof(int32_t));
    int32_t* Dm = (int32_t*)calloc(N, sizeof(int32_t));  /* queryDeletionMemo  (vertical gap) */
    int32_t* Im = (int32_t*)calloc(N, sizeof(int32_t));  /* queryInsertionMemo (horizontal gap) */
    uint8_t* T  = (uint8_t*)calloc(N, 1);
    const int32_t go = p->gap_open, ge = p->gap_extend;
    for (int i = 1; i <= Q; ++i) { H[i * W] = go + i * ge; T[i * W] = D_UP; }    /* :43-46 */
    for (int j = 1; j <= R; ++j) { H[j] = go + j * ge; T[j] = D_LEFT; }          /* :50-53 */
    for (int i = 1; i <= Q; ++i) {
        for (int j = 1; j <= R; ++j) {
            int pred; uint8_t t = 0;
            int32_t dv, iv;
            if (i == 1) { dv = H[(i - 1) * W + j] + go + ge; t |= 4; }           /* :185-189 */
            else { dv = vibmax_s32(H[(i - 1) * W + j] + go + ge, Dm[(i - 1) * W + j] + ge, &pred); if (pred) t |= 4; } /* :190-197 */
            if (j == 1) { iv = H[i * W + (j - 1)] + go + ge; t |= 8; }           /* :201-205 */
            else { iv = vibmax_s32(H[i * W + (j - 1)] + go + ge, Im[i * W + (j - 1)] + ge, &pred); if (pred) t |= 8; } /* :206-213 */
            Dm[i * W + j] = dv; Im[i * W + j] = iv;
            int32_t diag = H[(i - 1) * W + (j - 1)] + (q[i - 1] == r[j - 1] ? p->match : p->mismatch); /* :219-227 */
            uint8_t d = D_DIAG;
            int32_t m = vibmax_s32(dv, diag, &pred); if (pred) d = D_UP;         /* :229-230 */
            m = vibmax_s32(iv, m, &pred);            if (pred) d = D_LEFT;       /* :232-233 */
            H[i * W + j] = m; T[i * W + j] = (uint8_t)(t | d);
        }
    }
    *score = H[(size_t)Q * W + R];                                               /* :387 */
    int64_t len = -1;
    if (o_ref) {
        sb3 s; sb_init(&s, o_ref, o_rel, o_qry, (int64_t)Q + R);
        int i = Q, j = R; int state = 0;                  /* 0 SCORING, 1 INSERTION, 2 DELETION  (:253) */
        while (i != 0 && j != 0) {                                               /* :259 */
            uint8_t t = T[i * W + j];
            if (state == 0) {
                switch (t & 3) {
                    case D_DIAG: sb_push(&s, r[j - 1], q[i - 1] == r[j - 1] ? '*' : '|', q[i - 1]); --i; --j; break; /* :266-281 */
                    case D_UP:   state = 2; break;                               /* :284-290 */
                    case D_LEFT: state = 1; break;                               /* :293-299 */
                    default: abort();
                }
            } else if (state == 1) {                                             /* :309-332 */
                state = (t & 8) ? 0 : 1;
                sb_push(&s, r[j - 1], ' ', '_'); --j;
            } else {                                                             /* :334-357 */
                state = (t & 4) ? 0 : 2;
                sb_push(&s, '_', ' ', q[i - 1]); --i;
            }
        }
        while (i > 0) { sb_push(&s, '_', ' ', q[i - 1]); --i; }                  /* :366-371 */
        while (j > 0) { sb_push(&s, r[j - 1], ' ', '_'); --j; }                  /* :373-378 */
        len = sb_finish(&s);
    }
    free(H); free(Dm); free(Im); free(T);
    return len;
}

/* ---------------------------------------------------------------------------------
 * LinearSmithWaterman: score_matrix c++/LinearSmithWaterman.cpp:70-114, backtrack
 * :116-228 (BACKTRACK_ALL off), print_results :240-288.  Tie precedence UP > LEFT > DIAG.
 * band < 0 : unbanded.  band >= 0 : repaired BandedSmithWaterman semantics (header). */
static int64_t lsw_pair(const orc_params* p, int band, const char* r, int R, const char* q, int Q,
                        int32_t* score, int32_t* end_row, int32_t* end_col,
                        char* o_ref, char* o_rel, char* o_qry) {
    size_t W = (size_t)R + 1, N = W * ((size_t)Q + 1);
    int32_t* H = (int32_t*)calloc(N, sizeof(int32_t));
    uint8_t* D = (uint8_t*)calloc(N, 1);
    const int32_t g = p->gap_open;
    for (int i = 1; i <= Q; ++i) {
        int jlo = 1, jhi = R;
        if (band >= 0) { jlo = i - band < 1 ? 1 : i - band; jhi = (int64_t)i + band > R ? R : i + band; }
        for (int j = jlo; j <= jhi; ++j) {
            int32_t up = H[(i - 1) * W + j] + g;                                  /* :81 */
            int32_t left = H[i * W + (j - 1)] + g;                                /* :82 */
            int32_t diag = H[(i - 1) * W + (j - 1)] + (q[i - 1] == r[j - 1] ? p->match : p->mismatch); /* :86-94 */
            int32_t lc = left > diag ? left : diag;
            int32_t t = up > lc ? up : lc;                                        /* :97 */
            int32_t h = t > 0 ? t : 0;                                            /* :100 */
            H[i * W + j] = h;
            if (t < 0) continue;                                                  /* :103 */
            else if (up == h) D[i * W + j] = D_UP;                                /* :105 */
            else if (left == h) D[i * W + j] = D_LEFT;                            /* :106 */
            else D[i * W + j] = D_DIAG;                                           /* :107 */
        }
    }
    int32_t best = 0; int bi = 0, bj = 0;
    for (int i = 0; i <= Q; ++i)                                                  /* :145-157 first strict max, row-major */
        for (int j = 0; j <= R; ++j)
            if (H[i * W + j] > best) { best = H[i * W + j]; bi = i; bj = j; }
    *score = best; if (end_row) *end_row = bi; if (end_col) *end_col = bj;
    int64_t len = -1;
    if (o_ref) {
        sb3 s; sb_init(&s, o_ref, o_rel, o_qry, (int64_t)Q + R);
        if (best > 0) {                                   /* score 0: queue empty, three empty lines (:253-257) */
            int i = bi, j = bj;
            for (;;) {                                                            /* :163-226 */
                switch (D[i * W + j]) {
                    case D_DIAG: sb_push(&s, r[j - 1], q[i - 1] == r[j - 1] ? '*' : '|', q[i - 1]); --i; --j; break;
                    case D_LEFT: sb_push(&s, r[j - 1], ' ', '_'); --j; break;
                    case D_UP:   sb_push(&s, '_', ' ', q[i - 1]); --i; break;
                    default: abort();
                }
                if (H[i * W + j] == 0) break;                                     /* :222 */
            }
        }
        len = sb_finish(&s);
    }
    free(H); free(D);
    return len;
}

/* Band-only-memory restatement of the repaired banded SW (O(Q*(2W+1)) memory) so the
 * 10 kbp x 10 kbp, W=64 configuration can be checked on the host.  Identical results to
 * lsw_pair(band=W); pinned against it in tests/test_oracle.py.
 * Storage: row i keeps columns j in [i-W, i+W] at slot k = j - (i - W), 0 <= k <= 2W. */
static int64_t bsw_pair_bandmem(const orc_params* p, const char* r, int R, const char* q, int Q,
                                int32_t* score, int32_t* end_row, int32_t* end_col,
                                char* o_ref, char* o_rel, char* o_qry) {
    const int Wd = p->band; const size_t BW = (size_t)2 * Wd + 1;
    int32_t* H = (int32_t*)calloc(BW * ((size_t)Q + 1), sizeof(int32_t));
    uint8_t* D = (uint8_t*)calloc(BW * ((size_t)Q + 1), 1);
    const int32_t g = p->gap_open;
    #define HB(i, j) (((j) < (i) - Wd || (j) > (i) + Wd || (j) < 0 || (j) > R || (i) < 0) ? 0 : H[(size_t)(i) * BW + (size_t)((j) - ((i) - Wd))])
    int32_t best = 0; int bi = 0, bj = 0;
    for (int i = 1; i <= Q; ++i) {
        int jlo = i - Wd < 1 ? 1 : i - Wd, jhi = (int64_t)i + Wd > R ? R : i + Wd;
        for (int j = jlo; j <= jhi; ++j) {
            int32_t up = HB(i - 1, j) + g, left = HB(i, j - 1) + g;
            int32_t diag = HB(i - 1, j - 1) + (q[i - 1] == r[j - 1] ? p->match : p->mismatch);
            int32_t lc = left > diag ? left : diag, t = up > lc ? up : lc, h = t > 0 ? t : 0;
            size_t at = (size_t)i * BW + (size_t)(j - (i - Wd));
            H[at] = h;
            if (t >= 0) D[at] = (up == h) ? D_UP : (left == h) ? D_LEFT : D_DIAG;
            if (h > best) { best = h; bi = i; bj = j; }   /* row-major visiting order inside the band == global row-major order */
        }
    }
    *score = best; if (end_row) *end_row = bi; if (end_col) *end_col = bj;
    int64_t len = -1;
    if (o_ref) {
        sb3 s; sb_init(&s, o_ref, o_rel, o_qry, (int64_t)Q + R);
        if (best > 0) {
            int i = bi, j = bj;
            for (;;) {
                switch (D[(size_t)i * BW + (size_t)(j - (i - Wd))]) {
                    case D_DIAG: sb_push(&s, r[j - 1], q[i - 1] == r[j - 1] ? '*' : '|', q[i - 1]); --i; --j; break;
                    case D_LEFT: sb_push(&s, r[j - 1], ' ', '_'); --j; break;
                    case D_UP:   sb_push(&s, '_', ' ', q[i - 1]); --i; break;
                    default: abort();
                }
                if (HB(i, j) == 0) break;
            }
        }
        len = sb_finish(&s);
    }
    #undef HB
    free(H); free(D);
    return len;
}

/* ---------------------------------------------------------------------------------
 * Affine banded Smith-Waterman (ORC_ABSW) — see the header: local Gotoh on the cells with |i-j| <= band (band < 0: all
 * cells).  Border and out-of-band cells: H = 0 and D = I = "minus infinity" (a gap can neither start nor continue there).
 *   D[i][j] = vibmax(H[i-1][j] + open + extend, D[i-1][j] + extend)   tie -> GAP_OPEN   (c++/AffineNeedlemanWunsch.cpp:190-197)
 *   I[i][j] = vibmax(H[i][j-1] + open + extend, I[i][j-1] + extend)   tie -> GAP_OPEN   (:206-213)
 *   t = max(D, I, diag + s);  H = max(0, t);  dirH = UP if D == H, else LEFT if I == H, else DIAG   (c++/LinearSmithWaterman.cpp:97-108)
 * End cell: first strict maximum in row-major order (:145-157).  Walk: AffineNeedlemanWunsch's three states (:257-364) —
 * SCORING follows dirH (UP / LEFT switch state without moving), a gap state emits, moves and returns to SCORING when its
 * direction bit says GAP_OPEN — and stops like LinearSmithWaterman when the cell it arrives at in SCORING has H == 0 (:222). */
#define ORC_NEG (-(1 << 29))
static int64_t absw_pair(const orc_params* p, int band, const char* r, int R, const char* q, int Q,
                         int32_t* score, int32_t* end_row, int32_t* end_col,
                         char* o_ref, char* o_rel, char* o_qry) {
    size_t W = (size_t)R + 1, N = W * ((size_t)Q + 1);
    int32_t* H  = (int32_t*)calloc(N, sizeof(int32_t));
    int32_t* Dm = (int32_t*)malloc(N * sizeof(int32_t));
    int32_t* Im = (int32_t*)malloc(N * sizeof(int32_t));
    uint8_t* T  = (uint8_t*)calloc(N, 1);
    for (size_t k = 0; k < N; ++k) { Dm[k] = ORC_NEG; Im[k] = ORC_NEG; }
    const int32_t goe = p->gap_open + p->gap_extend, ge = p->gap_extend;
    for (int i = 1; i <= Q; ++i) {
        int jlo = 1, jhi = R;
        if (band >= 0) { jlo = i - band < 1 ? 1 : i - band; jhi = (int64_t)i + band > R ? R : i + band; }
        for (int j = jlo; j <= jhi; ++j) {
            int pred; uint8_t t = 0;
            int32_t dv = vibmax_s32(H[(i - 1) * W + j] + goe, Dm[(i - 1) * W + j] + ge, &pred); if (pred) t |= 4;
            int32_t iv = vibmax_s32(H[i * W + (j - 1)] + goe, Im[i * W + (j - 1)] + ge, &pred); if (pred) t |= 8;
            int32_t diag = H[(i - 1) * W + (j - 1)] + (q[i - 1] == r[j - 1] ? p->match : p->mismatch);
            int32_t m = dv > iv ? dv : iv; m = m > diag ? m : diag;
            int32_t h = m > 0 ? m : 0;
            Dm[i * W + j] = dv; Im[i * W + j] = iv; H[i * W + j] = h;
            if (h > 0) t |= (dv == h) ? D_UP : (iv == h) ? D_LEFT : D_DIAG;
            T[i * W + j] = t;
        }
    }
    int32_t best = 0; int bi = 0, bj = 0;
    for (int i = 0; i <= Q; ++i)
        for (int j = 0; j <= R; ++j)
            if (H[i * W + j] > best) { best = H[i * W + j]; bi = i; bj = j; }
    *score = best; if (end_row) *end_row = bi; if (end_col) *end_col = bj;
    int64_t len = -1;
    if (o_ref) {
        sb3 s; sb_init(&s, o_ref, o_rel, o_qry, (int64_t)Q + R);
        if (best > 0) {
            int i = bi, j = bj, state = 0;                 /* 0 SCORING, 1 INSERTION, 2 DELETION */
            for (;;) {
                const uint8_t t = T[i * W + j];
                if (state == 0) {
                    if (H[i * W + j] == 0) break;
                    switch (t & 3) {
                        case D_DIAG: sb_push(&s, r[j - 1], q[i - 1] == r[j - 1] ? '*' : '|', q[i - 1]); --i; --j; break;
                        case D_UP:   state = 2; break;
                        case D_LEFT: state = 1; break;
                        default: abort();
                    }
                } else if (state == 1) {
                    state = (t & 8) ? 0 : 1;
                    sb_push(&s, r[j - 1], ' ', '_'); --j;
                } else {
                    state = (t & 4) ? 0 : 2;
                    sb_push(&s, '_', ' ', q[i - 1]); --i;
                }
            }
        }
        len = sb_finish(&s);
    }
    free(H); free(Dm); free(Im); free(T);
    return len;
}

/* Linear-memory score + end-cell restatement of LinearSmithWaterman (a8/a9) for inputs
 * whose full matrix cannot be allocated (the reference needs 8 B/cell).  One rolling row.
 * Same first-strict-max-in-row-major rule (c++/LinearSmithWaterman.cpp:145-157).
 * band < 0 : unbanded. */
int orc_lsw_score_only(const orc_params* p, int band, const char* r, int64_t R, const char* q, int64_t Q,
                       int32_t* score, int64_t* end_row, int64_t* end_col) {
    /* row[j] holds H[i-1][j] until cell (i,j) overwrites it with H[i][j].
     * Banded case needs no explicit zeroing: the only out-of-band reads are
     *   H[i][jlo-1]   (left of the band's first cell)      -> `left` starts at 0,
     *   H[i-1][i+W]   (above the band's last cell)         -> never written by rows < i, still 0 from calloc,
     * and H[i-1][jlo-1] is in band for row i-1 (or column 0). */
    int32_t* row = (int32_t*)calloc((size_t)R + 1, sizeof(int32_t));
    if (!row) return -1;
    const int32_t g = p->gap_open, ma = p->match, mi = p->mismatch;
    int32_t best = 0; int64_t bi = 0, bj = 0;
    for (int64_t i = 1; i <= Q; ++i) {
        int64_t jlo = 1, jhi = R;
        if (band >= 0) { jlo = i - band < 1 ? 1 : i - band; jhi = i + band > R ? R : i + band; }
        const char qc = q[i - 1];
        int32_t diag = row[jlo - 1];
        int32_t left = 0;
        for (int64_t j = jlo; j <= jhi; ++j) {
            int32_t upv = row[j];
            int32_t d = diag + (qc == r[j - 1] ? ma : mi);
            int32_t u = upv + g, l = left + g;
            int32_t t = u > l ? u : l; t = t > d ? t : d;
            int32_t h = t > 0 ? t : 0;
            diag = upv; row[j] = h; left = h;
            if (h > best) { best = h; bi = i; bj = j; }
        }
    }
    *score = best; if (end_row) *end_row = bi; if (end_col) *end_col = bj;
    free(row);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * Public single-pair entry.  o_ref/o_rel/o_qry: NULL (scores only) or buffers of
 * Q+R+1 bytes each; returns alignment length (or -1 when strings not requested). */
int64_t orc_align_pair(const orc_params* p, const char* ref, int R, const char* qry, int Q,
                       int32_t* score, int32_t* end_row, int32_t* end_col,
                       char* o_ref, char* o_rel, char* o_qry) {
    switch (p->algo) {
        case ORC_LNW: if (end_row) *end_row = Q; if (end_col) *end_col = R;
                      return lnw_pair(p, ref, R, qry, Q, score, o_ref, o_rel, o_qry);
        case ORC_ANW: if (end_row) *end_row = Q; if (end_col) *end_col = R;
                      return anw_pair(p, ref, R, qry, Q, score, o_ref, o_rel, o_qry);
        case ORC_LSW: return lsw_pair(p, -1, ref, R, qry, Q, score, end_row, end_col, o_ref, o_rel, o_qry);
        case ORC_BSW: return lsw_pair(p, p->band, ref, R, qry, Q, score, end_row, end_col, o_ref, o_rel, o_qry);
        case ORC_ABSW: return absw_pair(p, p->band, ref, R, qry, Q, score, end_row, end_col, o_ref, o_rel, o_qry);
        default: return -2;
    }
}

int64_t orc_bsw_bandmem(const orc_params* p, const char* ref, int R, const char* qry, int Q,
                        int32_t* score, int32_t* end_row, int32_t* end_col,
                        char* o_ref, char* o_rel, char* o_qry) {
    return bsw_pair_bandmem(p, ref, R, qry, Q, score, end_row, end_col, o_ref, o_rel, o_qry);
}

/* ---------------------------------------------------------------------------------
 * Batch entry over a parseInput-style blob (c++/parseInput.cpp:78-113: NUL-separated
 * sequences + seqPair index).  Contiguous pair ranges per thread.
 * strings: NULL or a buffer with per-pair slots; slot i starts at str_off[i] and holds
 * 3 consecutive NUL-terminated strings REF, REL, QRY, each in a (Q+R+1)-byte field. */
typedef struct {
    const orc_params* p; const char* seqs; const orc_pair* pairs; size_t begin, end;
    int32_t* scores; int32_t* end_rc; char* strings; const int64_t* str_off; int bandmem;
} orc_job;

static void* orc_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    for (size_t i = j->begin; i < j->end; ++i) {
        const orc_pair* pr = &j->pairs[i];
        const char* r = j->seqs + pr->referenceIdx; const char* q = j->seqs + pr->queryIdx;
        int R = pr->referenceSize, Q = pr->querySize;
        char *o0 = NULL, *o1 = NULL, *o2 = NULL;
        if (j->strings) { size_t f = (size_t)Q + R + 1; o0 = j->strings + j->str_off[i]; o1 = o0 + f; o2 = o1 + f; }
        int32_t er = 0, ec = 0;
        if (j->bandmem && j->p->algo == ORC_BSW)
            bsw_pair_bandmem(j->p, r, R, q, Q, &j->scores[i], &er, &ec, o0, o1, o2);
        else
            orc_align_pair(j->p, r, R, q, Q, &j->scores[i], &er, &ec, o0, o1, o2);
        if (j->end_rc) { j->end_rc[2 * i] = er; j->end_rc[2 * i + 1] = ec; }
    }
    return NULL;
}

int orc_align_batch(const orc_params* p, const char* seqs, const orc_pair* pairs, size_t n, int threads,
                    int32_t* scores, int32_t* end_rc, char* strings, const int64_t* str_off, int bandmem) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n && n > 0) threads = (int)n;
    orc_job* jobs = (orc_job*)malloc(sizeof(orc_job) * (size_t)threads);
    pthread_t* tids = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (orc_job){p, seqs, pairs, n * (size_t)t / (size_t)threads, n * ((size_t)t + 1) / (size_t)threads,
                            scores, end_rc, strings, str_off, bandmem};
        if (threads == 1) orc_worker(&jobs[t]); else pthread_create(&tids[t], NULL, orc_worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(jobs); free(tids);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * LinearSmithWaterman with BACKTRACK_ALL (c++/LinearSmithWaterman.h:9, off as shipped; pinned on the reference compiled with
 * -DBACKTRACK_ALL, oracle/_ref/ref_align_all): EVERY cell holding the maximum starts a walk.  Start cells are queued bottom-right
 * to top-left (rows descending, columns descending, c++/LinearSmithWaterman.cpp:126-143); the queue advances every walk by one
 * move in turn (:163-226), so finished alignments come out ordered by their number of moves, ties in queue order.  Each walk
 * follows the ONE stored direction per cell (UP > LEFT > DIAG), like the single-path mode.  Score 0: the reference's all-maxima
 * mode walks uninitialised cells (undefined behaviour); this restatement prints the three empty lines of the single-path mode.
 * Returns a malloc'ed text of the reference's stdout blocks for the batch ("<i> | <score>" then REF / REL / QRY per alignment). */
typedef struct { int i, j; int64_t len; size_t order; } orc_start;
static int orc_start_cmp(const void* a, const void* b) {
    const orc_start* x = (const orc_start*)a; const orc_start* y = (const orc_start*)b;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}
int orc_lsw_all_text(const orc_params* p, const char* seqs, const orc_pair* pairs, size_t n, int first_index,
                     char** text_out, size_t* bytes_out, int64_t* n_alignments) {
    size_t cap = 1 << 16, len = 0; char* out = (char*)malloc(cap);
    int64_t total = 0;
    #define ORC_RESERVE(extra) do { if (len + (size_t)(extra) + 64 > cap) { while (len + (size_t)(extra) + 64 > cap) cap *= 2; out = (char*)realloc(out, cap); } } while (0)
    for (size_t k = 0; k < n; ++k) {
        const char* r = seqs + pairs[k].referenceIdx; const char* q = seqs + pairs[k].queryIdx;
        const int R = pairs[k].referenceSize, Q = pairs[k].querySize;
        const size_t W = (size_t)R + 1, N = W * ((size_t)Q + 1);
        int32_t* H = (int32_t*)calloc(N, sizeof(int32_t));
        uint8_t* D = (uint8_t*)calloc(N, 1);
        const int32_t g = p->gap_open;
        int32_t best = 0;
        for (int i = 1; i <= Q; ++i)
            for (int j = 1; j <= R; ++j) {
                int32_t up = H[(i - 1) * W + j] + g, left = H[i * W + (j - 1)] + g;
                int32_t diag = H[(i - 1) * W + (j - 1)] + (q[i - 1] == r[j - 1] ? p->match : p->mismatch);
                int32_t lc = left > diag ? left : diag, t = up > lc ? up : lc, h = t > 0 ? t : 0;
                H[i * W + j] = h;
                if (h > best) best = h;
                if (t < 0) continue;
                D[i * W + j] = up == h ? D_UP : (left == h ? D_LEFT : D_DIAG);
            }
        ORC_RESERVE(32);
        len += (size_t)snprintf(out + len, cap - len, "%d | %d\n", first_index + (int)k, best);
        if (best == 0) { ORC_RESERVE(3); memcpy(out + len, "\n\n\n", 3); len += 3; free(H); free(D); continue; }
        size_t ns = 0, scap = 16; orc_start* st = (orc_start*)malloc(scap * sizeof(orc_start));
        for (int i = Q; i >= 1; --i)
            for (int j = R; j >= 1; --j)
                if (H[i * W + j] == best) {
                    if (ns == scap) { scap *= 2; st = (orc_start*)realloc(st, scap * sizeof(orc_start)); }
                    int a = i, b = j; int64_t L = 0;
                    for (;;) {
                        const uint8_t d = D[a * W + b];
                        if (d == D_DIAG) { --a; --b; } else if (d == D_LEFT) --b; else --a;
                        ++L;
                        if (H[a * W + b] == 0) break;
                    }
                    st[ns] = (orc_start){i, j, L, ns}; ++ns;
                }
        qsort(st, ns, sizeof(orc_start), orc_start_cmp);
        for (size_t m = 0; m < ns; ++m) {
            const int64_t L = st[m].len;
            ORC_RESERVE(3 * (L + 1));
            char* o0 = out + len; char* o1 = o0 + L + 1; char* o2 = o1 + L + 1;
            int a = st[m].i, b = st[m].j; int64_t pos = L;
            while (pos > 0) {
                const uint8_t d = D[a * W + b];
                --pos;
                if (d == D_DIAG) { o0[pos] = r[b - 1]; o1[pos] = q[a - 1] == r[b - 1] ? '*' : '|'; o2[pos] = q[a - 1]; --a; --b; }
                else if (d == D_LEFT) { o0[pos] = r[b - 1]; o1[pos] = ' '; o2[pos] = '_'; --b; }
                else { o0[pos] = '_'; o1[pos] = ' '; o2[pos] = q[a - 1]; --a; }
            }
            o0[L] = o1[L] = o2[L] = '\n';
            len += (size_t)(3 * (L + 1));
        }
        total += (int64_t)ns;
        free(st); free(H); free(D);
    }
    #undef ORC_RESERVE
    out[len] = 0;
    *text_out = out; *bytes_out = len; if (n_alignments) *n_alignments = total;
    return 0;
}
void orc_free(void* p) { free(p); }

/* Reference stdout block for one pair: "<pairNum> | <score>\nREF\nREL\nQRY\n"
 * (c++/LinearNeedlemanWunsch.cpp:207-213, c++/AffineNeedlemanWunsch.cpp:386-391,
 * c++/LinearSmithWaterman.cpp:252-279; LSW score 0 prints three empty lines :253-257,
 * which is the same bytes as three empty strings).  Returns bytes written (excl. NUL). */
size_t orc_format_block(char* out, size_t cap, int pair_num, int32_t score,
                        const char* ref, const char* rel, const char* qry) {
    int n = snprintf(out, cap, "%d | %d\n%s\n%s\n%s\n", pair_num, score, ref, rel, qry);
    return n < 0 ? 0 : (size_t)n;
}
