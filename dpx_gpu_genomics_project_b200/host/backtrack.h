// Direction vocabulary of the reference (c++/backtrack.h:14-33), kept for source compatibility of callers.
// The traceback itself runs on the GPU (csrc/backtrack.cuh) over packed 2/4-bit codes.
#pragma once

enum directionMain { NONE_MAIN, MATCH, MISMATCH, QUERY_INSERTION, QUERY_DELETION };
enum directionIndel { NONE_INDEL, GAP_OPEN, GAP_EXTEND };
enum currentMatrixPosition { SCORING, INSERTION, DELETION };
