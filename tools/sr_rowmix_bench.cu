// sr_rowmix_bench.cu — how much of the DPX pipe can the short-read kernel's column step reach, as a function of the number of
// resident warps per SM sub-partition and of the number of INDEPENDENT row chains inside a step?
//
// The step of shortread.cuh (one column, K = 19 rows per lane) is a serial chain over the rows: h = vimax3(e, up, B) -> hg = h + G
// (IMAD) -> next row's `up`, about 10 clocks per row when nothing else delays it; PRMT, VIADDMNMX, the key IMAD and the best-key
// VIMNMX3 hang off it.  This benchmark replays exactly that instruction mix on registers (no memory), with the rows of a step split
// into CH chains (CH = 1: the kernel as it is; CH = 2: upper / lower half of the lane's rows as two columns one step apart), and
// reports clocks per step and the share of the 64-lane DPX pipe that the 3.5 ALU-pipe instructions per row keep busy.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sr_rowmix_bench tools/sr_rowmix_bench.cu
// usage: sr_rowmix_bench [out.json]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ROWS = 19;
constexpr int STEPS = 4096;

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}
__device__ __forceinline__ uint32_t mad(uint32_t a, uint32_t m, uint32_t b) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b)); return d;
}

__device__ __forceinline__ uint32_t add_alu(uint32_t a, uint32_t b) {                 // a plain add (IADD3 / VIADD unless ptxas moves it)
    uint32_t d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t maxu2_alu(uint32_t a, uint32_t b) {               // two-input packed max (VIMNMX.U16x2), kept apart from a second one
    uint32_t d; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}

// Variants of the per-row mix (V): bit 0: hg = h + G as a plain add instead of an IMAD; bit 1: key = (h << 7) + code as a shift-add (LEA)
// instead of an IMAD; bit 2: best-key update as two two-input maxima instead of one VIMNMX3.
// SHFL: the carried value of chain 0 goes through a shuffle, as in the kernel (24 clocks); TRACK: key + best-key update
template <int CH, bool SHFL, bool TRACK, int V = 0>
__global__ void __launch_bounds__(128, 4) rowmix(uint32_t seed, uint32_t one, uint32_t kmul, uint32_t* out, long long* t_first, long long* t_last) {
    uint32_t ta[ROWS], tb[ROWS], hg[ROWS], best[(ROWS + 1) / 2];
    #pragma unroll
    for (int r = 0; r < ROWS; ++r) { ta[r] = seed * (r + 3) + threadIdx.x; tb[r] = seed ^ (r * 0x9e3779b9u); hg[r] = 0x00020002u; }
    #pragma unroll
    for (int r = 0; r < (ROWS + 1) / 2; ++r) best[r] = 0;
    const uint32_t B2 = 0x00020002u, G2 = 0xfffdfffeu;
    uint32_t carried[CH], dcar[CH];
    #pragma unroll
    for (int c = 0; c < CH; ++c) { carried[c] = B2; dcar[c] = B2; }
    uint32_t rs = 0x5140u + (threadIdx.x & 3), cs = 0x003f003fu;
    constexpr int PER = (ROWS + CH - 1) / CH;
    const long long t0 = clock64();
    #pragma unroll 1
    for (int s = 0; s < STEPS; ++s) {
        uint32_t upg[CH], diag[CH], keyprev = 0;
        #pragma unroll
        for (int c = 0; c < CH; ++c) { upg[c] = carried[c]; diag[c] = dcar[c]; dcar[c] = carried[c]; }
        #pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int c = r / PER;
            const uint32_t sc = prmt_b32(ta[r], tb[r], rs);
            const uint32_t e = __viaddmax_s16x2(diag[c], sc, hg[r]);
            const uint32_t h = __vimax3_s16x2(e, upg[c], B2);
            diag[c] = hg[r];
            hg[r] = (V & 1) ? add_alu(h, G2) : mad(h, one, G2);
            upg[c] = hg[r];
            if (TRACK) {
                const uint32_t key = (V & 2) ? (h << 7) + cs : mad(h, kmul, cs);
                if (V & 4) best[r >> 1] = maxu2_alu(best[r >> 1], key);
                else if (r & 1) best[r >> 1] = __vimax3_u16x2(best[r >> 1], keyprev, key);
                else if (r == ROWS - 1) best[r >> 1] = __vmaxu2(best[r >> 1], key);
                keyprev = key;
            }
        }
        // chain c + 1 continues next step where chain c stopped this step; chain 0 takes the bottom row of the lane above
        #pragma unroll
        for (int c = CH - 1; c > 0; --c) carried[c] = upg[c - 1];
        carried[0] = SHFL ? __shfl_up_sync(0xffffffffu, upg[CH - 1], 1, 8) : upg[CH - 1];
        rs = (rs + 0x0101u) & 0x7373u;
        cs = mad(cs, one, 0xfffeffffu);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int r = 0; r < ROWS; ++r) acc ^= hg[r];
    #pragma unroll
    for (int r = 0; r < (ROWS + 1) / 2; ++r) acc ^= best[r];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) {
        unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicMin((unsigned long long*)&t_first[smid], (unsigned long long)t0);
        atomicMax((unsigned long long*)&t_last[smid], (unsigned long long)t1);
    }
}

template <int CH, bool SHFL, bool TRACK, int V = 0>
static void run(int sms, int warps_per_smsp, FILE* js, bool& first) {
    const int blocks = sms * warps_per_smsp;              // 128 threads per block = one warp per sub-partition and block
    uint32_t* out; long long *tf, *tl;
    CK(cudaMalloc(&out, (size_t)blocks * 128 * 4)); CK(cudaMalloc(&tf, 1024 * sizeof(long long))); CK(cudaMalloc(&tl, 1024 * sizeof(long long)));
    rowmix<CH, SHFL, TRACK, V><<<blocks, 128>>>(1234u, 1u, 128u, out, tf, tl);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(tf, 0x7f, 1024 * sizeof(long long))); CK(cudaMemset(tl, 0, 1024 * sizeof(long long)));
    rowmix<CH, SHFL, TRACK, V><<<blocks, 128>>>(4321u, 1u, 128u, out, tf, tl);
    CK(cudaDeviceSynchronize());
    std::vector<long long> hf(1024), hl(1024);
    CK(cudaMemcpy(hf.data(), tf, 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl.data(), tl, 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; int n = 0;
    for (int i = 0; i < 1024; ++i) if (hl[i] > 0 && hl[i] > hf[i]) { avg += (double)(hl[i] - hf[i]); ++n; }
    avg /= n ? n : 1;
    const double clk_per_step = avg / STEPS;                                   // all warps of a sub-partition advance one step
    const double alu = ROWS * (TRACK ? ((V & 4) ? 3.0 : 3.5) : 3.0) * 2.0 * warps_per_smsp;   // clocks of the half-rate (DPX, PRMT) instructions per step
    printf("chains %d  shfl %d  track %d  variant %d  warps/SMSP %d : %7.1f clk per step, %5.1f clk per row and warp, DPX pipe %5.1f %%\n",
           CH, (int)SHFL, (int)TRACK, V, warps_per_smsp, clk_per_step, clk_per_step / ROWS / warps_per_smsp, 100.0 * alu / clk_per_step);
    fprintf(js, "%s\n  {\"chains\": %d, \"shfl\": %d, \"track\": %d, \"variant\": %d, \"warps_per_smsp\": %d, \"clk_per_step\": %.2f, \"dpx_pipe_busy\": %.4f}",
            first ? "" : ",", CH, (int)SHFL, (int)TRACK, V, warps_per_smsp, clk_per_step, alu / clk_per_step);
    first = false;
    cudaFree(out); cudaFree(tf); cudaFree(tl);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    FILE* js = fopen(argc > 1 ? argv[1] : "sr_rowmix.json", "w");
    if (!js) { perror("open"); return 1; }
    fprintf(js, "{\"gpu\": \"%s\", \"sms\": %d, \"rows\": %d, \"results\": [", p.name, p.multiProcessorCount, ROWS);
    bool first = true;
    for (int w : {1, 2, 3, 4, 5, 6, 8}) {
        if (w > 5) continue;                                  // <= 96 registers here: five blocks of 128 threads per SM (the kernel: four)
        run<1, true, true>(p.multiProcessorCount, w, js, first);
        run<2, true, true>(p.multiProcessorCount, w, js, first);
        run<3, true, true>(p.multiProcessorCount, w, js, first);
        run<1, false, true>(p.multiProcessorCount, w, js, first);
        run<1, true, false>(p.multiProcessorCount, w, js, first);
        run<2, true, false>(p.multiProcessorCount, w, js, first);
        if (w != 4) continue;
        run<1, true, true, 1>(p.multiProcessorCount, w, js, first);
        run<1, true, true, 2>(p.multiProcessorCount, w, js, first);
        run<1, true, true, 3>(p.multiProcessorCount, w, js, first);
        run<1, true, true, 4>(p.multiProcessorCount, w, js, first);
        run<1, true, true, 5>(p.multiProcessorCount, w, js, first);
        run<1, true, true, 7>(p.multiProcessorCount, w, js, first);
        run<2, true, true, 3>(p.multiProcessorCount, w, js, first);
        run<2, true, true, 7>(p.multiProcessorCount, w, js, first);
        run<1, true, false, 1>(p.multiProcessorCount, w, js, first);
        run<2, true, false, 1>(p.multiProcessorCount, w, js, first);
    }
    fprintf(js, "\n]}\n");
    fclose(js);
    return 0;
}
