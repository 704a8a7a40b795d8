// SequenceAligner.h — source-compatible with the reference's abstract aligner interface
// (reference c++/SequenceAligner.h:6-28): same protected members, same constructor, same six virtuals.
// The subclasses in this directory are thin shims over the batched C ABI (include/dpxalign.h): the DP fill and
// the traceback run on the GPU; there is no CPU implementation behind them.
#pragma once
#include <string>

class SequenceAligner {
  protected:
    std::string reference_str;
    std::string query_str;
    int pairNum;

  public:
    SequenceAligner(const std::string input_reference, const std::string input_query, const int pairNum)
        : reference_str(input_reference), query_str(input_query), pairNum(pairNum) {}
    virtual ~SequenceAligner() {}

    virtual void init_matrix() = 0;
    virtual void print_matrix() = 0;
    virtual void score_matrix() = 0;
    virtual void backtrack() = 0;
    virtual void align() = 0;
    virtual void print_results() = 0;
};
