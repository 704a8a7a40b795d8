// Drop-in for reference c++/AffineNeedlemanWunsch.h:59-66 (same constructor argument order).
#pragma once
#include "GpuAligner.h"

class AffineNeedlemanWunsch : public dpxhost::GpuAligner {
  public:
    AffineNeedlemanWunsch(const std::string input_reference, const std::string input_query, const int pairNum,
                          const int matchWeight, const int mismatchWeight, const int gapOpenWeight, const int gapExtendWeight)
        : GpuAligner(input_reference, input_query, pairNum,
                     dpxhost::make_params(DPX_ALGO_ANW, matchWeight, mismatchWeight, gapOpenWeight, gapExtendWeight, 0)) {}
};
