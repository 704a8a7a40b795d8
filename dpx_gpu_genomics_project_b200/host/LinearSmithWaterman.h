// Drop-in for reference c++/LinearSmithWaterman.h:51-57 (same constructor argument order).
#pragma once
#include "GpuAligner.h"

class LinearSmithWaterman : public dpxhost::GpuAligner {
  public:
    LinearSmithWaterman(const std::string input_reference, const std::string input_query, const int pairNum,
                        const int match_weight, const int mismatch_weight, const int gap_weight)
        : GpuAligner(input_reference, input_query, pairNum,
                     dpxhost::make_params(DPX_ALGO_LSW, match_weight, mismatch_weight, gap_weight, 0, 0)) {}
};
