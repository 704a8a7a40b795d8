// dpx_ops.cuh — the hardware DPX instructions the reference only emulates on the CPU
// (c++/FakeDPX.{hpp,cpp}), exposed through one evaluation kernel so that the reference's own
// known-answer vectors (c++/testFakeDPX.cpp:10-113) can be replayed against the real sm_100a
// instructions (VIMNMX / VIMNMX3 / VIADDMNMX, .RELU, .S16x2/.U16x2, predicate-returning VIMNMX).
#pragma once
#include "common.cuh"

namespace dpx {

enum DpxOp : int {
    OP_VIMAX3_S32 = 0, OP_VIMAX3_S16X2, OP_VIMAX3_U32, OP_VIMAX3_U16X2,
    OP_VIMIN3_S32, OP_VIMIN3_S16X2, OP_VIMIN3_U32, OP_VIMIN3_U16X2,
    OP_VIMAX_S32_RELU, OP_VIMAX_S16X2_RELU, OP_VIMIN_S32_RELU, OP_VIMIN_S16X2_RELU,
    OP_VIMAX3_S32_RELU, OP_VIMAX3_S16X2_RELU, OP_VIMIN3_S32_RELU, OP_VIMIN3_S16X2_RELU,
    OP_VIBMAX_S32, OP_VIBMAX_U32, OP_VIBMIN_S32, OP_VIBMIN_U32,
    OP_VIBMAX_S16X2, OP_VIBMAX_U16X2, OP_VIBMIN_S16X2, OP_VIBMIN_U16X2,
    OP_VIADDMAX_S32, OP_VIADDMAX_U32, OP_VIADDMIN_S32, OP_VIADDMIN_U32,
    OP_VIADDMAX_S16X2, OP_VIADDMAX_U16X2, OP_VIADDMIN_S16X2, OP_VIADDMIN_U16X2,
    OP_VIADDMAX_S32_RELU, OP_VIADDMIN_S32_RELU, OP_VIADDMAX_S16X2_RELU, OP_VIADDMIN_S16X2_RELU,
    OP_COUNT
};

__global__ void dpx_eval_kernel(int op, const uint32_t* __restrict__ A, const uint32_t* __restrict__ B,
                                const uint32_t* __restrict__ Cc, int n, uint32_t* __restrict__ out,
                                uint8_t* __restrict__ phi, uint8_t* __restrict__ plo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t a = A[t], b = B[t], c = Cc[t];
    const int sa = (int)a, sb = (int)b, sc = (int)c;
    uint32_t r = 0; bool ph = false, pl = false;
    switch (op) {
        case OP_VIMAX3_S32:        r = (uint32_t)__vimax3_s32(sa, sb, sc); break;
        case OP_VIMAX3_S16X2:      r = __vimax3_s16x2(a, b, c); break;
        case OP_VIMAX3_U32:        r = __vimax3_u32(a, b, c); break;
        case OP_VIMAX3_U16X2:      r = __vimax3_u16x2(a, b, c); break;
        case OP_VIMIN3_S32:        r = (uint32_t)__vimin3_s32(sa, sb, sc); break;
        case OP_VIMIN3_S16X2:      r = __vimin3_s16x2(a, b, c); break;
        case OP_VIMIN3_U32:        r = __vimin3_u32(a, b, c); break;
        case OP_VIMIN3_U16X2:      r = __vimin3_u16x2(a, b, c); break;
        case OP_VIMAX_S32_RELU:    r = (uint32_t)__vimax_s32_relu(sa, sb); break;
        case OP_VIMAX_S16X2_RELU:  r = __vimax_s16x2_relu(a, b); break;
        case OP_VIMIN_S32_RELU:    r = (uint32_t)__vimin_s32_relu(sa, sb); break;
        case OP_VIMIN_S16X2_RELU:  r = __vimin_s16x2_relu(a, b); break;
        case OP_VIMAX3_S32_RELU:   r = (uint32_t)__vimax3_s32_relu(sa, sb, sc); break;
        case OP_VIMAX3_S16X2_RELU: r = __vimax3_s16x2_relu(a, b, c); break;
        case OP_VIMIN3_S32_RELU:   r = (uint32_t)__vimin3_s32_relu(sa, sb, sc); break;
        case OP_VIMIN3_S16X2_RELU: r = __vimin3_s16x2_relu(a, b, c); break;
        case OP_VIBMAX_S32:        r = (uint32_t)__vibmax_s32(sa, sb, &ph); break;
        case OP_VIBMAX_U32:        r = __vibmax_u32(a, b, &ph); break;
        case OP_VIBMIN_S32:        r = (uint32_t)__vibmin_s32(sa, sb, &ph); break;
        case OP_VIBMIN_U32:        r = __vibmin_u32(a, b, &ph); break;
        case OP_VIBMAX_S16X2:      r = __vibmax_s16x2(a, b, &ph, &pl); break;
        case OP_VIBMAX_U16X2:      r = __vibmax_u16x2(a, b, &ph, &pl); break;
        case OP_VIBMIN_S16X2:      r = __vibmin_s16x2(a, b, &ph, &pl); break;
        case OP_VIBMIN_U16X2:      r = __vibmin_u16x2(a, b, &ph, &pl); break;
        case OP_VIADDMAX_S32:      r = (uint32_t)__viaddmax_s32(sa, sb, sc); break;
        case OP_VIADDMAX_U32:      r = __viaddmax_u32(a, b, c); break;
        case OP_VIADDMIN_S32:      r = (uint32_t)__viaddmin_s32(sa, sb, sc); break;
        case OP_VIADDMIN_U32:      r = __viaddmin_u32(a, b, c); break;
        case OP_VIADDMAX_S16X2:    r = __viaddmax_s16x2(a, b, c); break;
        case OP_VIADDMAX_U16X2:    r = __viaddmax_u16x2(a, b, c); break;
        case OP_VIADDMIN_S16X2:    r = __viaddmin_s16x2(a, b, c); break;
        case OP_VIADDMIN_U16X2:    r = __viaddmin_u16x2(a, b, c); break;
        case OP_VIADDMAX_S32_RELU:   r = (uint32_t)__viaddmax_s32_relu(sa, sb, sc); break;
        case OP_VIADDMIN_S32_RELU:   r = (uint32_t)__viaddmin_s32_relu(sa, sb, sc); break;
        case OP_VIADDMAX_S16X2_RELU: r = __viaddmax_s16x2_relu(a, b, c); break;
        case OP_VIADDMIN_S16X2_RELU: r = __viaddmin_s16x2_relu(a, b, c); break;
        default: break;
    }
    out[t] = r; phi[t] = ph ? 1 : 0; plo[t] = pl ? 1 : 0;
}

}  // namespace dpx
