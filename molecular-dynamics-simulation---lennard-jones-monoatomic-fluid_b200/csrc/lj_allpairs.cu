// lj_allpairs.cu -- O(N^2) Lennard-Jones force / energy / virial kernels (FP64, sm_100a).
//
// Reference: /root/reference/scripts/physics/lj_potential_energy.f90:109-183 (the i<j loop) and
// physics/geometry_pbc.f90:86 (minimum image).  Bounded by the FP64 (DFMA) pipe: positions of
// even the largest all-pairs case (N=32k: 768 KB) live in L2/shared memory.
//
// Variant 0 ("full matrix"): the CTA (i-tile, j-split) owns `block` i-atoms (one per thread,
// in registers) and streams its j range through shared memory in tiles; every thread reads the
// same j (broadcast LDS).  Each ORDERED pair is evaluated, only a_i is accumulated, and the
// energy-like sums carry weight 1/2.  (i,j) and (j,i) produce bit-identical r2 and exactly
// opposite forces, so Newton's third law holds to the last bit.
//
// Determinism: no floating-point atomics.  Forces: per-split partials summed in split order by
// the finalize kernel.  Scalars: warp-shuffle tree -> per-CTA partial -> fixed-order final pass.
#include "ljmd_internal.h"
#include "lj_pair.cuh"
#include "reduce.cuh"

namespace ljmd {

struct PairArgs {
    SimConst sc;
    const double *rx, *ry, *rz; // [R][n]
    double *fpart;              // [R][n_splits][3][n]
    double *spart;              // [R][gridDim.x*gridDim.y][3]
    int split_len;
    unsigned long long *pair_counts; // optional [2]
};

template <int BLOCK, bool COUNT>
__global__ void __launch_bounds__(BLOCK) lj_pairs_full_kernel(PairArgs a)
{
    __shared__ double sx[BLOCK], sy[BLOCK], sz[BLOCK];
    __shared__ double scratch[3 * (BLOCK / 32)];

    const int n = a.sc.n;
    const int rep = blockIdx.z;
    const double *rx = a.rx + (size_t)rep * n, *ry = a.ry + (size_t)rep * n,
                 *rz = a.rz + (size_t)rep * n;
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;

    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < n;
    const double xi = valid ? rx[i] : 0.0, yi = valid ? ry[i] : 0.0, zi = valid ? rz[i] : 0.0;

    double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0, du = 0.0, ddu = 0.0;
    unsigned long long ncut = 0;

    const int j0 = blockIdx.y * a.split_len;
    const int j1 = min(n, j0 + a.split_len);
    for (int jt = j0; jt < j1; jt += BLOCK) {
        const int cnt = min(BLOCK, j1 - jt);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            sx[threadIdx.x] = rx[jt + threadIdx.x];
            sy[threadIdx.x] = ry[jt + threadIdx.x];
            sz[threadIdx.x] = rz[jt + threadIdx.x];
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const double dx = mic(xi - sx[jj], L, invL); // :119-126
            const double dy = mic(yi - sy[jj], L, invL);
            const double dz = mic(zi - sz[jj], L, invL);
            const double r2 = dist2(dx, dy, dz);         // :129
            if (r2 < rc2 && (jt + jj) != i) {            // :132 (self pair excluded: j /= i)
                const PairTerms t = pair_terms(dx, dy, dz, r2);
                fx += t.fx; fy += t.fy; fz += t.fz;      // :153-155
                u += t.u; du += t.du; ddu += t.ddu;      // :140,177,178
                if (COUNT) ++ncut;
            }
        }
    }

    if (valid) {
        double *fp = a.fpart + ((size_t)rep * gridDim.y + blockIdx.y) * 3 * n;
        fp[i] = fx; fp[(size_t)n + i] = fy; fp[2 * (size_t)n + i] = fz;
    } else {
        u = 0.0; du = 0.0; ddu = 0.0; ncut = 0;
    }
    double s[3] = {u, du, ddu};
    block_sum<3>(s, scratch);
    if (threadIdx.x == 0) {
        double *sp = a.spart +
                     ((size_t)rep * gridDim.x * gridDim.y + (size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
        sp[0] = s[0]; sp[1] = s[1]; sp[2] = s[2];
    }
    if (COUNT) {
        // integer counters: order-independent, so atomics are deterministic here
        ncut = __reduce_add_sync(0xffffffffu, (unsigned)ncut);
        if ((threadIdx.x & 31) == 0 && a.pair_counts) atomicAdd(&a.pair_counts[1], ncut);
    }
}

int pairwork_plan(ljmd_ctx *c)
{
    PairWork &w = c->pw;
    const int n = c->sc.n;
    w.variant = 0;
    w.block = 128;
    w.ipt = 1;
    w.n_itiles = (n + w.block * w.ipt - 1) / (w.block * w.ipt);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // aim for ~6 CTAs per SM across the whole launch (all replicas)
    const long long target = (long long)sms * 6;
    long long per_rep = (target + c->R - 1) / c->R;
    int splits = (int)((per_rep + w.n_itiles - 1) / w.n_itiles);
    if (splits < 1) splits = 1;
    int split_len = (n + splits - 1) / splits;
    split_len = ((split_len + 31) / 32) * 32;
    if (split_len < 32) split_len = 32;
    w.split_len = split_len;
    w.n_splits = (n + split_len - 1) / split_len;
    const size_t fbytes = sizeof(double) * (size_t)c->R * w.n_splits * 3 * n;
    const size_t sbytes = sizeof(double) * (size_t)c->R * w.n_itiles * w.n_splits * 3;
    LJMD_CUDA_CHECK(cudaMalloc(&w.fpart, fbytes));
    LJMD_CUDA_CHECK(cudaMalloc(&w.spart, sbytes));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.fpart, 0, fbytes, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.spart, 0, sbytes, c->stream));
    return LJMD_OK;
}

int launch_pairs(ljmd_ctx *c)
{
    PairArgs a;
    a.sc = c->sc;
    a.rx = c->st.rx; a.ry = c->st.ry; a.rz = c->st.rz;
    a.fpart = c->pw.fpart;
    a.spart = c->pw.spart;
    a.split_len = c->pw.split_len;
    a.pair_counts = c->fw.pair_counts;
    dim3 grid(c->pw.n_itiles, c->pw.n_splits, c->R);
    if (c->fw.pair_counts)
        lj_pairs_full_kernel<128, true><<<grid, 128, 0, c->stream>>>(a);
    else
        lj_pairs_full_kernel<128, false><<<grid, 128, 0, c->stream>>>(a);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

} // namespace ljmd
