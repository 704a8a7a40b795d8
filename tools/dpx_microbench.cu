// dpx_microbench.cu — measures the issue rate (lanes / clk / SM) of the integer / DPX instructions the
// alignment kernels are built from, on the GPU it runs on.  The score-only roofline of SURVEY.md §8d is
//     GCUPS_roof = lanes_per_clk_per_SM x SMs x f_clk / I_cell
// and `lanes_per_clk_per_SM` must be measured, not assumed.  Three views per instruction, all in SM CLOCKS read with clock64()
// (no wall-clock or frequency estimate enters a rate; v1 timed thread 0 of a block only and so under-counted the cycles):
//   throughput  8 independent chains per thread, 2048 threads per SM: lanes / clk / SM (cycles = last warp's end - first warp's start)
//   one warp    8 chains, ONE warp per SM sub-partition: what a single warp can issue (the long-pair chain runs like this)
//   latency     1 chain, one warp per SM: clocks between dependent issues
// The wrapper tools/run_microbench.py samples the NVML SM clock while this runs (profiles/r02_dpx_microbench.json).
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

enum Op { MAX_S32, VIMAX3_S32, VIADDMAX_S32, VIADDMAX_S32_RELU, VIMAX3_S32_RELU, VIBMAX_S32,
          VIMAX_S16X2, VIMAX3_S16X2, VIADDMAX_S16X2, VIADDMAX_S16X2_RELU, VIMAX3_S16X2_RELU, VIBMAX_S16X2_PRED,
          IADD, IADD3, LOP3, PRMT, IMAD, SHF, ISETP_SEL, SHFL_UP,
          MIX_VIADDMAX_IMAD, MIX_VIADDMAX16_IMAD, MIX_VIADDMAX_LOP3, MIX_PRMT_IMAD, MIX_2DPX_1IMAD, LDS32,
          CHAIN_RELU_IMAD, CHAIN_SHFL_RELU,
          MIX_DPX16_LOP3, MIX_DPX16_VIADD, MIX_DPX32_VIADD, MIX_DPX16_VIMNMX16, MIX_2DPX16_VIADD, MIX_PRMT_DPX16, VIMAX3_U16X2, MIX_DPX16_VIADD_IMAD,
          LOP3_VOL, MIX_DPX16_LOP3_VOL, MIX_DPX16_2LOP3_VOL, MIX_IMAD_LOP3_VOL, MIX_IMAD_VIADD, OP_COUNT };

static const char* op_name[OP_COUNT] = {
    "max.s32 (VIMNMX)", "__vimax3_s32 (VIMNMX3)", "__viaddmax_s32 (VIADDMNMX)", "__viaddmax_s32_relu (VIADDMNMX.RELU)",
    "__vimax3_s32_relu (VIMNMX3.RELU)", "__vibmax_s32 (ISETP+SEL/VIMNMX)",
    "__vimax_s16x2 (VIMNMX.S16x2)", "__vimax3_s16x2 (VIMNMX3.S16x2)", "__viaddmax_s16x2 (VIADDMNMX.S16x2)",
    "__viaddmax_s16x2_relu", "__vimax3_s16x2_relu", "__vibmax_s16x2 + 2 preds consumed",
    "IADD (a+b)", "IADD3 (a+b+c)", "LOP3 (a^b^c)", "PRMT (__byte_perm)", "IMAD (a*b+c)", "SHF (funnelshift)", "ISETP+SEL",
    "SHFL.UP", "mix: 1 VIADDMNMX + 1 IMAD", "mix: 1 VIADDMNMX.S16x2 + 1 IMAD", "mix: 1 VIADDMNMX + 1 LOP3",
    "mix: 1 PRMT + 1 IMAD", "mix: 2 VIADDMNMX + 1 IMAD", "LDS.32 (conflict-free)",
    "chain: VIADDMNMX.RELU -> IMAD (cross-pipe round trip)", "chain: SHFL.UP -> VIADDMNMX.RELU",
    "mix: 1 VIADDMNMX.S16x2 + 1 LOP3", "mix: 1 VIADDMNMX.S16x2 + 1 VIADD", "mix: 1 VIADDMNMX + 1 VIADD", "mix: 1 VIADDMNMX.S16x2 + 1 VIMNMX.S16x2",
    "mix: 2 VIADDMNMX.S16x2 + 1 VIADD", "mix: 1 PRMT + 1 VIADDMNMX.S16x2", "__vimax3_u16x2 (VIMNMX3.U16x2)", "mix: 1 VIADDMNMX.S16x2 + 1 VIADD + 1 IMAD",
    "LOP3 (asm volatile, majority)", "mix: 1 VIADDMNMX.S16x2 + 1 LOP3 (volatile)", "mix: 1 VIADDMNMX.S16x2 + 2 LOP3 (volatile)",
    "mix: 1 IMAD + 1 LOP3 (volatile)", "mix: 1 IMAD + 1 VIADD" };
static const int op_results_per_iter[OP_COUNT] = {1,1,1,1,1,1, 1,1,1,1,1,1, 1,1,1,1,1,1,1, 1, 2,2,2,2,3, 1, 2, 2, 2,2,2,2,3,2,1,3, 1,2,3,2,2};

template <int OP, int CH>
__global__ void __launch_bounds__(1024) bench(uint32_t seed, uint32_t* out, long long* t_first, long long* t_last) {
    __shared__ uint32_t sm[1024];
    uint32_t x[CH], y[CH];
    const uint32_t a = seed * 3 + threadIdx.x, b = seed ^ 0x9e3779b9u, c = (seed >> 3) | 1u;
    #pragma unroll
    for (int k = 0; k < CH; ++k) { x[k] = a + k * 7919u; y[k] = b + k; }
    sm[threadIdx.x] = a;
    __syncthreads();
    const long long t0 = clock64();
    constexpr int REP = CH == 1 ? 16 : 2;            // dependent ops per chain between two loop branches (the branch costs ~20 clocks)
    #pragma unroll 1
    for (int it = 0; it < ITERS / REP; ++it) {
        #pragma unroll
        for (int kk = 0; kk < CH * REP; ++kk) {
            const int k = kk % CH;
            if (OP == MAX_S32)            asm volatile("max.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));      // volatile: a repeated max would fold
            if (OP == VIMAX3_S32)         x[k] = (uint32_t)__vimax3_s32((int)x[k], (int)y[k], (int)c);
            if (OP == VIADDMAX_S32)       x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)y[k]);
            if (OP == VIADDMAX_S32_RELU)  x[k] = (uint32_t)__viaddmax_s32_relu((int)x[k], (int)c, (int)y[k]);
            if (OP == VIMAX3_S32_RELU)    x[k] = (uint32_t)__vimax3_s32_relu((int)x[k], (int)y[k], (int)c);
            if (OP == VIBMAX_S32)       { bool p; x[k] = (uint32_t)__vibmax_s32((int)x[k], (int)y[k], &p); y[k] += p ? 1u : 0u; }
            if (OP == VIMAX_S16X2)        asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
            if (OP == VIMAX3_S16X2)       x[k] = __vimax3_s16x2(x[k], y[k], c);
            if (OP == VIADDMAX_S16X2)     x[k] = __viaddmax_s16x2(x[k], c, y[k]);
            if (OP == VIADDMAX_S16X2_RELU) x[k] = __viaddmax_s16x2_relu(x[k], c, y[k]);
            if (OP == VIMAX3_S16X2_RELU)  x[k] = __vimax3_s16x2_relu(x[k], y[k], c);
            if (OP == VIBMAX_S16X2_PRED) { bool ph, pl; x[k] = __vibmax_s16x2(x[k], y[k], &ph, &pl); if (ph) y[k] = c; if (pl) y[k] ^= 1u; }
            if (OP == IADD)               asm volatile("add.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
            if (OP == IADD3)              asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(x[k]) : "r"(y[k]), "r"(c));
            if (OP == LOP3)               asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(c));
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
            if (OP == CHAIN_RELU_IMAD)    { x[k] = (uint32_t)__viaddmax_s32_relu((int)x[k], (int)c, (int)y[k]); x[k] = x[k] * c + b; }
            if (OP == MIX_DPX16_LOP3)     { x[k] = __viaddmax_s16x2(x[k], c, b); y[k] = (y[k] ^ a) & c; }
            if (OP == MIX_DPX16_VIADD)    { x[k] = __viaddmax_s16x2(x[k], c, b); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
            if (OP == MIX_DPX32_VIADD)    { x[k] = (uint32_t)__viaddmax_s32((int)x[k], (int)c, (int)b); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
            if (OP == MIX_DPX16_VIMNMX16) { x[k] = __viaddmax_s16x2(x[k], c, b); asm volatile("max.s16x2 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
            if (OP == MIX_2DPX16_VIADD)   { x[k] = __viaddmax_s16x2(x[k], c, b); x[k] = __vimax3_s16x2(x[k], a, c); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
            if (OP == MIX_PRMT_DPX16)     { x[k] = __viaddmax_s16x2(x[k], c, b); y[k] = __byte_perm(y[k], b, c); }
            if (OP == VIMAX3_U16X2)       x[k] = __vimax3_u16x2(x[k], y[k], c);
            if (OP == MIX_DPX16_VIADD_IMAD){ x[k] = __viaddmax_s16x2(x[k], c, b); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); y[k] = y[k] * c + b; }
            if (OP == LOP3_VOL)           asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(x[k]) : "r"(y[k]), "r"(c));
            if (OP == MIX_DPX16_LOP3_VOL) { x[k] = __viaddmax_s16x2(x[k], c, b); asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[k]) : "r"(a), "r"(c)); }
            if (OP == MIX_DPX16_2LOP3_VOL){ x[k] = __viaddmax_s16x2(x[k], c, b); asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[k]) : "r"(a), "r"(c));
                                            asm volatile("lop3.b32 %0, %0, %1, %2, 0xd4;" : "+r"(y[k]) : "r"(b), "r"(c)); }
            if (OP == MIX_IMAD_LOP3_VOL)  { x[k] = x[k] * c + b; asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[k]) : "r"(a), "r"(c)); }
            if (OP == MIX_IMAD_VIADD)     { x[k] = x[k] * c + b; asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
            if (OP == CHAIN_SHFL_RELU)    { x[k] = __shfl_up_sync(0xffffffffu, x[k], 1); x[k] = (uint32_t)__viaddmax_s32_relu((int)x[k], (int)c, (int)y[k]); }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int k = 0; k < CH; ++k) acc ^= x[k] + y[k];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    // clock64 is per SM: the first warp's start and the last warp's end of every block, reduced per SM on the host
    if ((threadIdx.x & 31) == 0) {
        unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicMin((unsigned long long*)&t_first[smid], (unsigned long long)t0);
        atomicMax((unsigned long long*)&t_last[smid], (unsigned long long)t1);
    }
}

struct Shape { const char* name; int blocks_per_sm, threads; };

template <int OP, int CH>
static double run_shape(int sms, const Shape& sh, float* ms_out) {
    const int blocks = sms * sh.blocks_per_sm;
    uint32_t* out; long long *tf, *tl;
    CK(cudaMalloc(&out, (size_t)blocks * sh.threads * 4)); CK(cudaMalloc(&tf, 1024 * sizeof(long long))); CK(cudaMalloc(&tl, 1024 * sizeof(long long)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<OP, CH><<<blocks, sh.threads>>>(1234u, out, tf, tl);     // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(tf, 0x7f, 1024 * sizeof(long long))); CK(cudaMemset(tl, 0, 1024 * sizeof(long long)));
    CK(cudaEventRecord(e0));
    bench<OP, CH><<<blocks, sh.threads>>>(4321u, out, tf, tl);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(ms_out, e0, e1));
    std::vector<long long> hf(1024), hl(1024);
    CK(cudaMemcpy(hf.data(), tf, 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl.data(), tl, 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; int n = 0;
    for (int i = 0; i < 1024; ++i) if (hl[i] > 0 && hl[i] > hf[i]) { avg += (double)(hl[i] - hf[i]); ++n; }
    cudaFree(out); cudaFree(tf); cudaFree(tl);
    return n ? avg / n : 0.0;                                    // SM clocks from the first warp's start to the last warp's end
}

template <int OP>
static void run_one(int sms, FILE* js, bool first) {
    static const Shape full{"full", 2, 1024}, onewarp{"one warp per SMSP", 1, 128}, single{"single warp", 1, 32};
    float ms = 0, ms2 = 0, ms3 = 0;
    const double cyc_full = run_shape<OP, 8>(sms, full, &ms);
    const double cyc_1w = run_shape<OP, 8>(sms, onewarp, &ms2);
    const double cyc_lat = run_shape<OP, 1>(sms, single, &ms3);
    const int rpi = op_results_per_iter[OP];
    const double per_clk = 2048.0 * ITERS * 8 * rpi / cyc_full;              // lanes / clk / SM at full occupancy
    const double per_clk_1w = 128.0 * ITERS * 8 * rpi / cyc_1w;              // lanes / clk / SM with one warp per sub-partition
    const double lat = cyc_lat / ((double)ITERS * rpi);                      // clocks per dependent instruction
    const double mhz = cyc_full / (ms * 1e3);
    printf("%-52s %7.2f lanes/clk/SM | 1 warp/SMSP %7.2f | dependent issue %6.2f clk | (%.3f ms, %.0f MHz implied)\n", op_name[OP], per_clk, per_clk_1w, lat, ms, mhz);
    fprintf(js, "%s\n  {\"op\": \"%s\", \"lanes_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm_one_warp_per_smsp\": %.3f, \"dependent_issue_clocks\": %.3f, \"ms\": %.4f, \"sm_mhz_implied\": %.0f}",
            first ? "" : ",", op_name[OP], per_clk, per_clk_1w, lat, ms, mhz);
}

template <int OP> struct Runner { static void go(int sms, FILE* js) { run_one<OP>(sms, js, OP == 0); Runner<OP + 1>::go(sms, js); } };
template <> struct Runner<OP_COUNT> { static void go(int, FILE*) {} };

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    FILE* js = fopen(argc > 1 ? argv[1] : "dpx_microbench.json", "w");
    if (!js) { perror("open"); return 1; }
    fprintf(js, "{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"method\": \"clock64 per SM: first warp start to last warp end; 8 chains (throughput), 1 chain (latency)\", \"results\": [",
            p.name, p.multiProcessorCount, p.clockRate);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    Runner<0>::go(p.multiProcessorCount, js);
    fprintf(js, "\n]}\n");
    fclose(js);
    return 0;
}
