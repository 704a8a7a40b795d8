// GpuAligner.h — shared implementation of the four reference aligner classes on top of libdpxalign.
// One pair = a one-element batch through dpx_align_batch (the per-pair align() of the reference is far too
// fine-grained for a GPU; the drop-in main.cpp submits whole files instead).
#pragma once
#include <cstdint>
#include <string>
#include "SequenceAligner.h"
#include "../../include/dpxalign.h"

namespace dpxhost {

// Process-wide device context (device 0 unless DPX_DEVICE is set).  Exits with a message when no GPU is
// available, the way the reference exits on its fatal errors (e.g. c++/parseInput.cpp:14).
dpx_ctx* engine();

class GpuAligner : public SequenceAligner {
  protected:
    dpx_params params;
    bool done = false;
    int32_t score = 0;
    int32_t end_row = 0, end_col = 0;
    std::string reference_sequence, pair_relation, query_sequence;

    void run();                       // fill + traceback on the GPU (idempotent)

  public:
    GpuAligner(const std::string& ref, const std::string& qry, int pairNum, const dpx_params& p)
        : SequenceAligner(ref, qry, pairNum), params(p) {}

    void init_matrix() override {}    // matrices live in registers / HBM of the device
    void print_matrix() override;     // not available: the score matrix is never materialised
    void score_matrix() override { run(); }
    void backtrack() override { run(); }
    void print_results() override;    // "<pairNum> | <score>\nREF\nREL\nQRY\n"
    void align() override { init_matrix(); score_matrix(); backtrack(); print_results(); }

    int getScore() { run(); return score; }
    int getEndRow() { run(); return end_row; }
    int getEndCol() { run(); return end_col; }
    const std::string& getReferenceSequence() { run(); return reference_sequence; }
    const std::string& getPairRelation() { run(); return pair_relation; }
    const std::string& getQuerySequence() { run(); return query_sequence; }
};

inline dpx_params make_params(int algo, int match, int mismatch, int gap_open, int gap_extend, int band) {
    dpx_params p;
    p.algo = algo; p.match = match; p.mismatch = mismatch; p.gap_open = gap_open; p.gap_extend = gap_extend; p.band = band;
    p.flags = DPX_OUT_SCORE | DPX_OUT_END_COORDS | DPX_OUT_STRINGS;
    return p;
}

}  // namespace dpxhost
