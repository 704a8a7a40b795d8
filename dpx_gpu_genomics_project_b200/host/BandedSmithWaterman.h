// Drop-in for reference c++/BandedSmithWaterman.h:51-57: pairNum is the LAST of the reference's arguments.
// The reference never initialises band_width (.h:16); here it is a trailing defaulted argument plus a setter.
// Semantics are the repaired ones documented in DESIGN.md (LinearSmithWaterman restricted to |i-j| <= band).
#pragma once
#include "GpuAligner.h"

class BandedSmithWaterman : public dpxhost::GpuAligner {
  public:
    BandedSmithWaterman(const std::string input_reference, const std::string input_query, const int match_weight,
                        const int mismatch_weight, const int gap_weight, const int pairNum, const int band_width = 64)
        : GpuAligner(input_reference, input_query, pairNum,
                     dpxhost::make_params(DPX_ALGO_BSW, match_weight, mismatch_weight, gap_weight, 0, band_width)) {}
    void set_band_width(int w) { params.band = w; done = false; }
};
