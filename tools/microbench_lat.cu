// microbench_lat.cu -- dependent-issue latency of FP64 ops on sm_100a and pipe utilisation as a
// function of resident warps x independent chains (the numbers that size the pair kernel).
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int CHAINS, int OP>
__global__ void k(double *out, double a, double b, int iters, long long *cyc)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c * 1e-3 + threadIdx.x * 1e-6;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (OP == 0) x[c] = fma(x[c], a, b);
                if (OP == 1) x[c] = __dadd_rn(x[c], b);
                if (OP == 2) x[c] = __dmul_rn(x[c], a);
                if (OP == 3) x[c] = rint(x[c] * a);             // DMUL + FRND
                if (OP == 4) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[c])); x[c] = r; }
                if (OP == 5) x[c] = __shfl_sync(0xffffffffu, x[c], (threadIdx.x + 1) & 31);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS, int OP>
int run(const char *name, int warps_per_sm, double *d, long long *dc)
{
    const int iters = 512;
    k<CHAINS, OP><<<148, warps_per_sm * 32>>>(d, 1.0000001, 1e-9, iters, dc);
    CK(cudaDeviceSynchronize());
    k<CHAINS, OP><<<148, warps_per_sm * 32>>>(d, 1.0000001, 1e-9, iters, dc);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    const double per_instr = (double)c / (iters * 16.0 * CHAINS);
    const double ops_per_clk_smsp = (warps_per_sm / 4.0) / per_instr;
    printf("%-10s chains=%d warps/SM=%2d : %7.2f cyc per instr per warp  => %.3f warp-instr/clk/SMSP (peak 0.5)\n", name, CHAINS,
           warps_per_sm, per_instr, ops_per_clk_smsp);
    return 0;
}

int main()
{
    double *d; long long *dc;
    CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&dc, 64));
    run<1, 0>("DFMA", 4, d, dc); run<2, 0>("DFMA", 4, d, dc); run<4, 0>("DFMA", 4, d, dc); run<8, 0>("DFMA", 4, d, dc);
    run<1, 1>("DADD", 4, d, dc); run<1, 2>("DMUL", 4, d, dc);
    run<1, 3>("DMUL+FRND", 4, d, dc); run<1, 4>("RCP64H", 4, d, dc); run<1, 5>("SHFL64", 4, d, dc);
    run<4, 0>("DFMA", 8, d, dc); run<4, 0>("DFMA", 16, d, dc); run<2, 0>("DFMA", 16, d, dc); run<1, 0>("DFMA", 16, d, dc);
    run<1, 0>("DFMA", 32, d, dc); run<2, 0>("DFMA", 32, d, dc);
    return 0;
}
