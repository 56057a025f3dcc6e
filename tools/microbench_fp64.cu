// microbench_fp64.cu -- issue-rate microbenchmarks that drive the design of the force kernels.
// Measures warp-instructions per clock per SM for the FP64 pipe and the side pipes the pair
// loop uses (FRND.F64, MUFU.RCP64H, SHFL, LDS.64, integer ALU), alone and mixed with DFMA.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench microbench_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;

enum Op { DADD, DMUL, DFMA, DSETP, FRND_RINT, FRND_FLOOR, ROUND_AWAY, RCP64H, DIV, SHFL64, LDS64_DISTINCT,
          LDS64_BCAST, IALU, MIX_DFMA_FRND, MIX_DFMA_SHFL, MIX_DFMA_LDS, MIX_DFMA_IALU, MIX_DFMA2_FRND1,
          MIX_DFMA4_FRND1, FFMA, MIX_DFMA_FFMA, NOPS };
const char *names[] = {"DADD", "DMUL", "DFMA", "DSETP+sel", "FRND.F64 rint", "FRND.F64 floor", "round() half-away",
                       "MUFU.RCP64H", "1.0/x IEEE", "SHFL x2 (64-bit)", "LDS.64 lane-distinct", "LDS.64 broadcast",
                       "IADD/LOP (int ALU)", "DFMA+FRND 1:1", "DFMA+SHFL64 1:1", "DFMA+LDS.64 1:1", "DFMA+IALU 1:1",
                       "DFMA+FRND 2:1", "DFMA+FRND 4:1", "FFMA", "DFMA+FFMA 1:1"};
// counted instructions per inner unroll item, for normalisation
const double counted[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 5, 1, 2};

template <int OP>
__global__ void __launch_bounds__(256) k(double *out, double a, double b, int n)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = threadIdx.x * 1e-3;
    __syncthreads();
    double x[UNROLL];
    int ix[UNROLL];
    float fx[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { x[u] = a + u * 1e-3 + threadIdx.x * 1e-6; ix[u] = threadIdx.x + u; fx[u] = (float)x[u]; }
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (OP == DADD) x[u] = __dadd_rn(x[u], b);
            if (OP == DMUL) x[u] = __dmul_rn(x[u], a);
            if (OP == DFMA) x[u] = fma(x[u], a, b);
            if (OP == DSETP) x[u] = (x[u] < b) ? a : x[u];
            if (OP == FRND_RINT) x[u] = rint(x[u]);
            if (OP == FRND_FLOOR) x[u] = floor(x[u]);
            if (OP == ROUND_AWAY) x[u] = round(x[u]);
            if (OP == RCP64H) { int hi = __double2hiint(x[u]); double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[u])); x[u] = r; (void)hi; }
            if (OP == DIV) x[u] = 1.0 / x[u];
            if (OP == SHFL64) x[u] = __shfl_sync(0xffffffffu, x[u], (lane + 1) & 31);
            if (OP == LDS64_DISTINCT) { ix[u] = (ix[u] + 1) & 255; x[u] = sm[ix[u]] ; ix[u] += __double2loint(x[u]) & 1; }
            if (OP == LDS64_BCAST) { ix[u] = (i + u) & 255; x[u] += sm[ix[u]]; }
            if (OP == IALU) ix[u] = (ix[u] ^ 0x5a5a) + i;
            if (OP == MIX_DFMA_FRND) { x[u] = fma(x[u], a, b); x[u] = rint(x[u]); }
            if (OP == MIX_DFMA_SHFL) { x[u] = fma(x[u], a, b); x[u] = __shfl_sync(0xffffffffu, x[u], (lane + 1) & 31); }
            if (OP == MIX_DFMA_LDS) { x[u] = fma(x[u], a, sm[(ix[u] + i) & 255]); }
            if (OP == MIX_DFMA_IALU) { x[u] = fma(x[u], a, b); ix[u] = (ix[u] ^ 0x5a5a) + i; }
            if (OP == MIX_DFMA2_FRND1) { x[u] = fma(x[u], a, b); x[u] = fma(x[u], a, b); x[u] = rint(x[u]); }
            if (OP == MIX_DFMA4_FRND1) { x[u] = fma(x[u], a, b); x[u] = fma(x[u], a, b); x[u] = fma(x[u], a, b); x[u] = fma(x[u], a, b); x[u] = rint(x[u]); }
            if (OP == FFMA) fx[u] = fmaf(fx[u], (float)a, (float)b);
            if (OP == MIX_DFMA_FFMA) { x[u] = fma(x[u], a, b); fx[u] = fmaf(fx[u], (float)a, (float)b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += x[u] + ix[u] + fx[u];
    if (s == 1.2345) out[0] = s;
}

template <int OP>
void run(double *d, int sms, double clk_ghz)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = sms * 8, block = 256;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        k<OP><<<grid, block>>>(d, 1.0000001, 1e-9, ITERS);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    const double warp_instr = (double)grid * (block / 32) * ITERS * UNROLL * counted[OP];
    const double per_s = warp_instr / (best * 1e-3);
    const double per_clk_sm = per_s / (clk_ghz * 1e9) / sms;
    printf("%-24s %8.3f ms  %7.3f warp-instr/clk/SM (at %.3f GHz)  => %6.2f cyc/warp-instr/SMSP\n", names[OP], best,
           per_clk_sm, clk_ghz, 4.0 / per_clk_sm);
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz * 1e-6;
    printf("device %s, %d SMs, clock attr %.3f GHz (rates assume this clock)\n", p.name, p.multiProcessorCount, ghz);
    double *d; CK(cudaMalloc(&d, 64));
    const int sms = p.multiProcessorCount;
    run<DFMA>(d, sms, ghz);  // warm clocks
    run<DADD>(d, sms, ghz); run<DMUL>(d, sms, ghz); run<DFMA>(d, sms, ghz); run<DSETP>(d, sms, ghz);
    run<FRND_RINT>(d, sms, ghz); run<FRND_FLOOR>(d, sms, ghz); run<ROUND_AWAY>(d, sms, ghz);
    run<RCP64H>(d, sms, ghz); run<DIV>(d, sms, ghz); run<SHFL64>(d, sms, ghz);
    run<LDS64_DISTINCT>(d, sms, ghz); run<LDS64_BCAST>(d, sms, ghz); run<IALU>(d, sms, ghz);
    run<MIX_DFMA_FRND>(d, sms, ghz); run<MIX_DFMA_SHFL>(d, sms, ghz); run<MIX_DFMA_LDS>(d, sms, ghz);
    run<MIX_DFMA_IALU>(d, sms, ghz); run<MIX_DFMA2_FRND1>(d, sms, ghz); run<MIX_DFMA4_FRND1>(d, sms, ghz);
    run<FFMA>(d, sms, ghz); run<MIX_DFMA_FFMA>(d, sms, ghz);
    return 0;
}
