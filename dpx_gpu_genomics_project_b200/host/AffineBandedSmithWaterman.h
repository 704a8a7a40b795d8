// AffineBandedSmithWaterman — NOT a reference class: the variant the reference only names as a TODO
// (python/LinearBandedSmithWaterman.py:8, "BSW evidently usually uses affine gap penalties"), shaped like the reference's
// BandedSmithWaterman constructor (c++/BandedSmithWaterman.h:51-57: weights first, pairNum after them) with the two gap weights
// of AffineNeedlemanWunsch (c++/AffineNeedlemanWunsch.h:59-66).  Semantics: include/dpxalign.h, DPX_ALGO_ABSW.
#pragma once
#include "GpuAligner.h"

class AffineBandedSmithWaterman : public dpxhost::GpuAligner {
  public:
    AffineBandedSmithWaterman(const std::string input_reference, const std::string input_query, const int match_weight,
                              const int mismatch_weight, const int gap_open_weight, const int gap_extend_weight, const int pairNum,
                              const int band_width = 64)
        : GpuAligner(input_reference, input_query, pairNum,
                     dpxhost::make_params(DPX_ALGO_ABSW, match_weight, mismatch_weight, gap_open_weight, gap_extend_weight, band_width)) {}
    void set_band_width(int w) { params.band = w; done = false; }
};
