// dpx_microbench.cu — measures the issue rate (lanes / clk / SM) of the integer / DPX instructions the
// alignment kernels are built from, on the GPU it runs on.  The score-only roofline of SURVEY.md §8d is
//     GCUPS_roof = lanes_per_clk_per_SM x SMs x f_clk / I_cell
// and `lanes_per_clk_per_SM` must be measured, not assumed.  Each test runs CH independent dependency
// chains per thread (ILP) at full occupancy and reports warp-lane results per SM clock, from clock64().
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dpx_microbench.cu -o tools/dpx_microbench
// run  : tools/dpx_microbench [out.json]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;          // independent chains per thread

enum Op { MAX_S32, VIMAX3_S32, VIADDMAX_S32, VIADDMAX_S32_RELU, VIMAX3_S32_RELU, VIBMAX_S32,
          VIMAX_S16X2, VIMAX3_S16X2, VIADDMAX_S16X2, VIADDMAX_S16X2_RELU, VIMAX3_S16X2_RELU, VIBMAX_S16X2_PRED,
          IADD, IADD3, LOP3, PRMT, IMAD, SHF, ISETP_SEL, SHFL_UP,
          MIX_VIADDMAX_IMAD, MIX_VIADDMAX16_IMAD, MIX_VIADDMAX_LOP3, MIX_PRMT_IMAD, MIX_2DPX_1IMAD, LDS32, OP_COUNT };

static const char* op_name[OP_COUNT] = {
    "max.s32 (VIMNMX)", "__vimax3_s32 (VIMNMX3)", "__viaddmax_s32 (VIADDMNMX)", "__viaddmax_s32_relu (VIADDMNMX.RELU)",
    "__vimax3_s32_relu (VIMNMX3.RELU)", "__vibmax_s32 (ISETP+SEL/VIMNMX)",
    "__vimax_s16x2 (VIMNMX.S16x2)", "__vimax3_s16x2 (VIMNMX3.S16x2)", "__viaddmax_s16x2 (VIADDMNMX.S16x2)",
    "__viaddmax_s16x2_relu", "__vimax3_s16x2_relu", "__vibmax_s16x2 + 2 preds consumed",
    "IADD (a+b)", "IADD3 (a+b+c)", "LOP3 (a^b&c)", "PRMT (__byte_perm)", "IMAD (a*b+c)", "SHF (funnelshift)", "ISETP+SEL",
    "SHFL.UP", "mix: 1 VIADDMNMX + 1 IMAD", "mix: 1 VIADDMNMX.S16x2 + 1 IMAD", "mix: 1 VIADDMNMX + 1 LOP3",
    "mix: 1 PRMT + 1 IMAD", "mix: 2 VIADDMNMX + 1 IMAD", "LDS.32 (conflict-free)" };
static const int op_results_per_iter[OP_COUNT] = {1,1,1,1,1,1, 1,1,1,1,1,1, 1,1,1,1,1,1,1, 1, 2,2,2,2,3, 1};

template <int OP>
__global__ void __launch_bounds__(1024) bench(uint32_t seed, uint32_t* out, long long* cycles) {
    __shared__ uint32_t sm[1024];
    uint32_t x[CH], y[CH];
    const uint32_t a = seed * 3 + threadIdx.x, b = seed ^ 0x9e3779b9u, c = (seed >> 3) | 1u;
    #pragma unroll
    for (int k = 0; k < CH; ++k) { x[k] = a + k * 7919u; y[k] = b + k; }
    sm[threadIdx.x] = a;
    __syncthreads();
    const long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (OP == MAX_S32)            x[k] = (uint32_t)max((int)x[k] + 0, (int)(b ^ k)) ^ 0, x[k] = (uint32_t)max((int)x[k], (int)y[k]);
            if (OP == VIMAX3_S32)         x[k] = (uint32_t)__vimax3_s32((int)x[k], (int)y[k], (int)c);
            if (OP == VIADDMAX_S32)       x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)y[k]);
            if (OP == VIADDMAX_S32_RELU)  x[k] = (uint32_t)__viaddmax_s32_relu((int)x[k], (int)c, (int)y[k]);
            if (OP == VIMAX3_S32_RELU)    x[k] = (uint32_t)__vimax3_s32_relu((int)x[k], (int)y[k], (int)c);
            if (OP == VIBMAX_S32)       { bool p; x[k] = (uint32_t)__vibmax_s32((int)x[k], (int)y[k], &p); y[k] += p ? 1u : 0u; }
            if (OP == VIMAX_S16X2)        x[k] = __vmaxs2(x[k], y[k]) + 0u, y[k] = y[k];
            if (OP == VIMAX3_S16X2)       x[k] = __vimax3_s16x2(x[k], y[k], c);
            if (OP == VIADDMAX_S16X2)     x[k] = __viaddmax_s16x2(x[k], c, y[k]);
            if (OP == VIADDMAX_S16X2_RELU) x[k] = __viaddmax_s16x2_relu(x[k], c, y[k]);
            if (OP == VIMAX3_S16X2_RELU)  x[k] = __vimax3_s16x2_relu(x[k], y[k], c);
            if (OP == VIBMAX_S16X2_PRED) { bool ph, pl; x[k] = __vibmax_s16x2(x[k], y[k], &ph, &pl); if (ph) y[k] = c; if (pl) y[k] ^= 1u; }
            if (OP == IADD)               x[k] = x[k] + y[k];
            if (OP == IADD3)              x[k] = x[k] + y[k] + c;
            if (OP == LOP3)               x[k] = (x[k] ^ y[k]) & (c + k);
            if (OP == PRMT)               x[k] = __byte_perm(x[k], y[k], c);
            if (OP == IMAD)               x[k] = x[k] * c + y[k];
            if (OP == SHF)                x[k] = __funnelshift_l(x[k], y[k], 7);
            if (OP == ISETP_SEL)          x[k] = ((int)x[k] > (int)y[k]) ? x[k] + 0u : c, x[k] ^= 0u;
            if (OP == SHFL_UP)            x[k] = __shfl_up_sync(0xffffffffu, x[k], 1);
            if (OP == MIX_VIADDMAX_IMAD)  { x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)b); y[k] = y[k] * c + a; }
            if (OP == MIX_VIADDMAX16_IMAD){ x[k] = __viaddmax_s16x2(x[k], c, b); y[k] = y[k] * c + a; }
            if (OP == MIX_VIADDMAX_LOP3)  { x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)b); y[k] = (y[k] ^ a) & c; }
            if (OP == MIX_PRMT_IMAD)      { x[k] = __byte_perm(x[k], b, c); y[k] = y[k] * c + a; }
            if (OP == MIX_2DPX_1IMAD)     { x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)b); x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)a, (int)c); y[k] = y[k] * c + a; }
            if (OP == LDS32)              x[k] = sm[(x[k] + threadIdx.x) & 1023u] ;
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int k = 0; k < CH; ++k) acc ^= x[k] + y[k];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_one(int sms, FILE* js, bool first) {
    const int blocks = sms * 2, threads = 1024;       // 2048 threads / SM = full occupancy
    uint32_t* out; long long* cyc;
    CK(cudaMalloc(&out, (size_t)blocks * threads * 4)); CK(cudaMalloc(&cyc, blocks * sizeof(long long)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<OP><<<blocks, threads>>>(1234u, out, cyc);     // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    bench<OP><<<blocks, threads>>>(4321u, out, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(blocks);
    CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= blocks;
    const double results_per_sm = 2048.0 * ITERS * CH * op_results_per_iter[OP];
    const double per_clk = results_per_sm / avg;         // both blocks of an SM run concurrently for ~avg cycles
    const double mhz = avg / (ms * 1e3);
    printf("%-46s %8.2f lanes/clk/SM   (%.3f ms, ~%.0f MHz)\n", op_name[OP], per_clk, ms, mhz);
    fprintf(js, "%s\n  {\"op\": \"%s\", \"lanes_per_clk_per_sm\": %.3f, \"ms\": %.4f, \"sm_mhz_est\": %.0f}", first ? "" : ",", op_name[OP], per_clk, ms, mhz);
    cudaFree(out); cudaFree(cyc);
}

template <int OP> struct Runner { static void go(int sms, FILE* js) { run_one<OP>(sms, js, OP == 0); Runner<OP + 1>::go(sms, js); } };
template <> struct Runner<OP_COUNT> { static void go(int, FILE*) {} };

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    FILE* js = fopen(argc > 1 ? argv[1] : "dpx_microbench.json", "w");
    if (!js) { perror("open"); return 1; }
    fprintf(js, "{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"chains_per_thread\": %d, \"threads_per_sm\": 2048, \"results\": [",
            p.name, p.multiProcessorCount, p.clockRate, CH);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    Runner<0>::go(p.multiProcessorCount, js);
    fprintf(js, "\n]}\n");
    fclose(js);
    return 0;
}
