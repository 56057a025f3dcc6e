// lj_allpairs.cu -- O(N^2) Lennard-Jones force / energy / virial kernels (FP64, sm_100a).
//
// Reference: /root/reference/scripts/physics/lj_potential_energy.f90:109-183 (the i<j loop) and
// physics/geometry_pbc.f90:86 (minimum image).  Bounded by the FP64 (DFMA) pipe: positions of
// even the largest all-pairs case (N=32k: 768 KB) live in L2/shared memory.
//
// Variant 1 ("symmetric", default): every UNORDERED pair i<j is evaluated exactly once, like the
// reference, and the force is applied to both atoms (Newton's third law).
//   * Work item = (row of RI*32*W i-atoms, run of 32-atom j-chunks at or above the row's
//     diagonal).  One CTA of W warps per work item; heavy rows are launched first.
//   * Each lane keeps RI i-atoms (position + force accumulator) in registers.  The 32 j-atoms of
//     a chunk sit in shared memory; at rotation step s lane l meets j-atom (l+s)&31, so all 32
//     lanes always work on different j-atoms.  The j-side force accumulators live in registers
//     and rotate one lane per step (SHFL); after 32 steps they are back in their home lane.
//   * j-side forces of the W warps are combined through shared memory in warp order and written
//     to a per-row partial buffer; i-side forces are written once per work item.  The finalize
//     kernel adds the partials of each atom in a fixed order => bit-reproducible, no FP atomics.
//   * Measured issue costs on B200 (tools/microbench_fp64.cu): DADD/DMUL/DFMA 2 cyc/warp/SMSP,
//     FRND.F64 9 cyc on a separate pipe, IEEE divide ~17 cyc, SHFL.64 8 cyc, LDS.64 8-10 cyc.
//     Hence: rint() for the minimum image (runs beside the FP64 pipe), a 3-DFMA reciprocal
//     (MUFU.RCP64H seed + one cubic step, <= 1 ulp) instead of the IEEE divide, and force
//     accumulation fused into DFMAs.
//   * The cutoff decision uses r2 computed in the reference's exact operation order, so the set
//     of interacting pairs is identical to the reference's.  Per-pair terms then differ from the
//     reference's by <= ~1 ulp (reciprocal, fused accumulation), far inside the 1e-12 budget.
//
// Variant 0 ("full matrix", kept as the simple cross-check kernel): each thread owns one i-atom
// and visits every j (ordered pairs, weight 1/2 on the energy sums), per-pair arithmetic
// bit-identical to the reference.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ljmd_internal.h"
#include "lj_pair.cuh"
#include "reduce.cuh"

namespace ljmd {

// ==========================================================================================
// Variant 0: full matrix
// ==========================================================================================
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
        const unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)ncut);
        if ((threadIdx.x & 31) == 0 && a.pair_counts) atomicAdd(&a.pair_counts[1], (unsigned long long)c);
    }
}

// ==========================================================================================
// Variant 1: symmetric tiles
// ==========================================================================================
struct SymArgs {
    SimConst sc;
    const double *rx, *ry, *rz; // [R][n]
    const int4 *items;          // work items: {row, first chunk, number of chunks, split index}
    double *fI;                 // [R][max_splits][3][n]  i-side partial forces
    double *fJ;                 // [R][n_rows][3][n]      j-side partial forces
    double *spart;              // [R][n_items][3]        (S12, S6, unused)
    int max_splits, n_rows;
    unsigned long long *pair_counts; // optional [2]
};

// 1/x to <= 1 ulp: MUFU.RCP64H seed (relative error < 2^-20) and one cubic (Halley) step,
// y = y0*(1 + e + e^2), e = 1 - x*y0  => error ~ e^3 < 2^-60 before the final rounding.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0, 1.0);
    const double t = fma(e, e, e);
    return fma(y0, t, y0);
}

// One 32-atom j-chunk against the RI*32 i-atoms of a warp: 32 rotation steps.
// The distance stage of the RI pairs of a step is computed first (RI independent dependency
// chains), then one branch-free interaction body handles all of them: a pair outside the cutoff
// contributes exact zeros (its 1/r^2 is replaced by 0), so no per-pair branch serialises the
// chains.  The body is skipped only when no pair of the step is inside the cutoff.
template <int RI, bool DIAG, bool COUNT>
__device__ __forceinline__ void sym_chunk(const double (*sjb)[32], const double (&xi)[RI], const double (&yi)[RI],
                                          const double (&zi)[RI], double (&fxi)[RI], double (&fyi)[RI],
                                          double (&fzi)[RI], double &fjx, double &fjy, double &fjz, double &s12,
                                          double &s6, unsigned &ncut, double L, double invL, double rc2, int lane,
                                          int iw0, int j0, int n)
{
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {
        const int jj = (lane + s) & 31;
        const double xj = sjb[0][jj], yj = sjb[1][jj], zj = sjb[2][jj];
        double dx[RI], dy[RI], dz[RI], r2[RI];
        bool in[RI];
        bool any = false;
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            dx[r] = mic(xi[r] - xj, L, invL); // lj_potential_energy.f90:119-126
            dy[r] = mic(yi[r] - yj, L, invL);
            dz[r] = mic(zi[r] - zj, L, invL);
            r2[r] = dist2(dx[r], dy[r], dz[r]); // :129, reference operation order
            in[r] = r2[r] < rc2;                // :132
            if (DIAG) {
                const int ig = iw0 + r * 32 + lane, jg = j0 + jj;
                in[r] = in[r] && (jg > ig) && (jg < n);
            }
            any = any || in[r];
        }
        if (any) {
#pragma unroll
            for (int r = 0; r < RI; ++r) {
                double ir2 = fast_rcp(in[r] ? r2[r] : 1.0);
                ir2 = in[r] ? ir2 : 0.0;
                const double ir6 = __dmul_rn(__dmul_rn(ir2, ir2), ir2);
                const double ir12 = __dmul_rn(ir6, ir6);
                s12 = __dadd_rn(s12, ir12);
                s6 = __dadd_rn(s6, ir6);
                const double g = __dmul_rn(fma(2.0, ir12, -ir6), ir2); // -dU_r / r^2  (:143,148)
                fxi[r] = fma(g, dx[r], fxi[r]); fyi[r] = fma(g, dy[r], fyi[r]); fzi[r] = fma(g, dz[r], fzi[r]);
                fjx = fma(-g, dx[r], fjx); fjy = fma(-g, dy[r], fjy); fjz = fma(-g, dz[r], fjz);
                if (COUNT) ncut += in[r] ? 1u : 0u;
            }
        }
        fjx = __shfl_sync(0xffffffffu, fjx, (lane + 1) & 31);
        fjy = __shfl_sync(0xffffffffu, fjy, (lane + 1) & 31);
        fjz = __shfl_sync(0xffffffffu, fjz, (lane + 1) & 31);
    }
}

template <int RI, int W, bool COUNT>
__global__ void __launch_bounds__(W * 32) lj_pairs_sym_kernel(SymArgs a)
{
    constexpr int TI = RI * 32; // i-atoms per warp
    __shared__ double sj[2][3][32];        // double-buffered positions of the j-chunk
    __shared__ double fjw[2][W][3][32];    // double-buffered per-warp j-side forces
    __shared__ double scratch[3 * W];

    const int n = a.sc.n;
    const int rep = blockIdx.y;
    const double *rx = a.rx + (size_t)rep * n, *ry = a.ry + (size_t)rep * n,
                 *rz = a.rz + (size_t)rep * n;
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;
    const int4 item = a.items[blockIdx.x];
    const int row = item.x, chunk0 = item.y, nchunks = item.z, split = item.w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int iw0 = (row * W + warp) * TI; // first i-atom of this warp

    double xi[RI], yi[RI], zi[RI], fxi[RI], fyi[RI], fzi[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
        const int ig = iw0 + r * 32 + lane;
        const bool v = ig < n;
        xi[r] = v ? rx[ig] : 0.0; yi[r] = v ? ry[ig] : 0.0; zi[r] = v ? rz[ig] : 0.0;
        fxi[r] = 0.0; fyi[r] = 0.0; fzi[r] = 0.0;
    }
    double s12 = 0.0, s6 = 0.0;
    unsigned ncut = 0;

    constexpr int NT = W * 32;
    constexpr int NST = (96 + NT - 1) / NT; // staging elements per thread (96 = 3 x 32 doubles)
    // stage the first chunk
#pragma unroll
    for (int q = 0; q < NST; ++q) {
        const int t = threadIdx.x + q * NT;
        if (t < 96) {
            const int c = t >> 5, jg = chunk0 * 32 + (t & 31);
            const double *src = (c == 0) ? rx : (c == 1) ? ry : rz;
            sj[0][c][t & 31] = (jg < n) ? src[jg] : 0.0;
        }
    }
    __syncthreads();

    for (int k = 0; k < nchunks; ++k) {
        const int buf = k & 1;
        const int j0 = (chunk0 + k) * 32;
        // prefetch the next chunk into a register (stored to the other buffer after the compute)
        double pre[NST];
#pragma unroll
        for (int q = 0; q < NST; ++q) {
            const int t = threadIdx.x + q * NT;
            pre[q] = 0.0;
            if (t < 96 && k + 1 < nchunks) {
                const int c = t >> 5, jg = j0 + 32 + (t & 31);
                const double *src = (c == 0) ? rx : (c == 1) ? ry : rz;
                if (jg < n) pre[q] = src[jg];
            }
        }

        double fjx = 0.0, fjy = 0.0, fjz = 0.0;
        // Which (i,j) of this warp x chunk are wanted?  j > i and j < n.
        //   j0 >= iw0 + TI and the chunk is full : all        (fast path, no predicate)
        //   j0 + 32 <= iw0                        : none       (warp idles for this chunk)
        //   otherwise                             : per pair   (diagonal / ragged chunk)
        const bool none = (j0 + 32 <= iw0);
        const bool all = (j0 >= iw0 + TI) && (j0 + 32 <= n);
        if (!none) {
            if (all)
                sym_chunk<RI, false, COUNT>(sj[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut,
                                            L, invL, rc2, lane, iw0, j0, n);
            else
                sym_chunk<RI, true, COUNT>(sj[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut,
                                           L, invL, rc2, lane, iw0, j0, n);
        }
        // after 32 rotations lane l again holds the accumulator of j-atom j0 + l
        fjw[buf][warp][0][lane] = fjx;
        fjw[buf][warp][1][lane] = fjy;
        fjw[buf][warp][2][lane] = fjz;
#pragma unroll
        for (int q = 0; q < NST; ++q) {
            const int t = threadIdx.x + q * NT;
            if (t < 96 && k + 1 < nchunks) sj[buf ^ 1][t >> 5][t & 31] = pre[q];
        }
        __syncthreads();
        // combine the W warps in warp order and write the row's j-side partial for this chunk
#pragma unroll
        for (int q = 0; q < NST; ++q) {
            const int t = threadIdx.x + q * NT;
            if (t < 96) {
                const int c = t >> 5, jl = t & 31, jg = j0 + jl;
                double f = fjw[buf][0][c][jl];
#pragma unroll
                for (int w = 1; w < W; ++w) f += fjw[buf][w][c][jl];
                if (jg < n) a.fJ[(((size_t)rep * a.n_rows + row) * 3 + c) * n + jg] = f;
            }
        }
    }

    // i-side partial forces of this work item
#pragma unroll
    for (int r = 0; r < RI; ++r) {
        const int ig = iw0 + r * 32 + lane;
        if (ig < n) {
            double *fp = a.fI + ((size_t)rep * a.max_splits + split) * 3 * n;
            fp[ig] = fxi[r]; fp[(size_t)n + ig] = fyi[r]; fp[2 * (size_t)n + ig] = fzi[r];
        }
    }
    double s[3] = {s12, s6, 0.0};
    block_sum<3>(s, scratch);
    if (threadIdx.x == 0) {
        double *sp = a.spart + ((size_t)rep * gridDim.x + blockIdx.x) * 3;
        sp[0] = s[0]; sp[1] = s[1]; sp[2] = 0.0;
    }
    if (COUNT) {
        const unsigned c = __reduce_add_sync(0xffffffffu, ncut);
        if (lane == 0 && a.pair_counts) atomicAdd(&a.pair_counts[1], (unsigned long long)c);
    }
}

// ------------------------------------------------------------------------------------------
// Planning
// ------------------------------------------------------------------------------------------
static int sym_chunks_per_item(int n, int tcta, int R, int sms)
{
    // total chunk visits over all rows (triangle), then pick the run length so that the launch
    // has about 3 CTAs per SM (all replicas together) but no run shorter than 2 chunks.
    const int nchunks = (n + 31) / 32;
    const int nrows = (n + tcta - 1) / tcta;
    long long total = 0;
    for (int r = 0; r < nrows; ++r) total += nchunks - (long long)r * tcta / 32;
    const long long target = std::max<long long>(1, (long long)sms * 3 / std::max(1, R));
    long long c = (total + target - 1) / target;
    if (c < 2) c = 2;
    if (c > nchunks) c = nchunks;
    return (int)c;
}

int pairwork_plan(ljmd_ctx *c)
{
    PairWork &w = c->pw;
    const int n = c->sc.n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    const char *force_variant = getenv("LJMD_ALLPAIRS_VARIANT");
    w.variant = 1;
    if (force_variant && force_variant[0] == '0') w.variant = 0;

    if (w.variant == 0) {
        w.block = 128;
        w.ipt = 1;
        w.n_itiles = (n + w.block - 1) / w.block;
        const long long target = (long long)sms * 6;
        long long per_rep = (target + c->R - 1) / c->R;
        int splits = (int)((per_rep + w.n_itiles - 1) / w.n_itiles);
        if (splits < 1) splits = 1;
        int split_len = (n + splits - 1) / splits;
        split_len = ((split_len + 31) / 32) * 32;
        if (split_len < 32) split_len = 32;
        w.split_len = split_len;
        w.n_splits = (n + split_len - 1) / split_len;
        w.n_rows = 0;
        w.n_items = w.n_itiles * w.n_splits;
        const size_t fbytes = sizeof(double) * (size_t)c->R * w.n_splits * 3 * n;
        LJMD_CUDA_CHECK(cudaMalloc(&w.fpart, fbytes));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.fpart, 0, fbytes, c->stream));
    } else {
        // geometry: large systems use 4 warps x 4 i-atoms per lane (512-atom rows); small ones
        // 2 x 2 (128-atom rows) so that there are enough work items.
        if (n >= 2048) { w.ri = 4; w.warps = 4; } else { w.ri = 2; w.warps = 2; }
        // tuning overrides (benchmark sweeps only)
        if (const char *e = getenv("LJMD_SYM_RI")) w.ri = atoi(e);
        if (const char *e = getenv("LJMD_SYM_W")) w.warps = atoi(e);
        if (!((w.ri == 4 && w.warps == 4) || (w.ri == 2 && w.warps == 4) || (w.ri == 4 && w.warps == 2) ||
              (w.ri == 2 && w.warps == 2) || (w.ri == 1 && w.warps == 4) || (w.ri == 4 && w.warps == 8) ||
              (w.ri == 2 && w.warps == 8) || (w.ri == 2 && w.warps == 1) || (w.ri == 4 && w.warps == 1)))
            return set_error(LJMD_ERR_ARG, "unsupported symmetric-kernel geometry RI=%d W=%d", w.ri, w.warps);
        const int tcta = w.ri * 32 * w.warps;
        const int nchunks = (n + 31) / 32;
        w.n_rows = (n + tcta - 1) / tcta;
        w.chunks_per_item = sym_chunks_per_item(n, tcta, c->R, sms);
        if (const char *e = getenv("LJMD_SYM_CPI")) w.chunks_per_item = std::max(1, std::min(nchunks, atoi(e)));
        std::vector<int4> items;
        int max_splits = 0;
        for (int r = 0; r < w.n_rows; ++r) {
            const int first = r * tcta / 32;
            int split = 0;
            for (int c0 = first; c0 < nchunks; c0 += w.chunks_per_item, ++split) {
                int4 it;
                it.x = r; it.y = c0; it.z = std::min(w.chunks_per_item, nchunks - c0); it.w = split;
                items.push_back(it);
            }
            max_splits = std::max(max_splits, split);
        }
        // heavy items first: full-length runs before the short tails (stable => deterministic)
        std::stable_sort(items.begin(), items.end(), [](const int4 &x, const int4 &y) { return x.z > y.z; });
        w.n_items = (int)items.size();
        w.n_splits = max_splits;
        LJMD_CUDA_CHECK(cudaMalloc(&w.items, sizeof(int4) * items.size()));
        LJMD_CUDA_CHECK(cudaMemcpy(w.items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice));
        const size_t ibytes = sizeof(double) * (size_t)c->R * max_splits * 3 * n;
        const size_t jbytes = sizeof(double) * (size_t)c->R * w.n_rows * 3 * n;
        LJMD_CUDA_CHECK(cudaMalloc(&w.fpart, ibytes));
        LJMD_CUDA_CHECK(cudaMalloc(&w.fJ, jbytes));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.fpart, 0, ibytes, c->stream));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.fJ, 0, jbytes, c->stream));
    }
    const size_t sbytes = sizeof(double) * (size_t)c->R * w.n_items * 3;
    LJMD_CUDA_CHECK(cudaMalloc(&w.spart, sbytes));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.spart, 0, sbytes, c->stream));
    return LJMD_OK;
}

template <int RI, int W>
static void launch_sym(ljmd_ctx *c, const SymArgs &a)
{
    dim3 grid(c->pw.n_items, c->R);
    if (a.pair_counts) lj_pairs_sym_kernel<RI, W, true><<<grid, W * 32, 0, c->stream>>>(a);
    else lj_pairs_sym_kernel<RI, W, false><<<grid, W * 32, 0, c->stream>>>(a);
}

int launch_pairs(ljmd_ctx *c, int buf)
{
    const double *px = c->rbuf[buf], *py = px + c->nR, *pz = py + c->nR;
    const PairWork &w = c->pw;
    if (w.variant == 0) {
        PairArgs a;
        a.sc = c->sc;
        a.rx = px; a.ry = py; a.rz = pz;
        a.fpart = w.fpart;
        a.spart = w.spart;
        a.split_len = w.split_len;
        a.pair_counts = c->fw.pair_counts;
        dim3 grid(w.n_itiles, w.n_splits, c->R);
        if (c->fw.pair_counts) lj_pairs_full_kernel<128, true><<<grid, 128, 0, c->stream>>>(a);
        else lj_pairs_full_kernel<128, false><<<grid, 128, 0, c->stream>>>(a);
    } else {
        SymArgs a;
        a.sc = c->sc;
        a.rx = px; a.ry = py; a.rz = pz;
        a.items = w.items;
        a.fI = w.fpart;
        a.fJ = w.fJ;
        a.spart = w.spart;
        a.max_splits = w.n_splits;
        a.n_rows = w.n_rows;
        a.pair_counts = c->fw.pair_counts;
        if (w.ri == 4 && w.warps == 4) launch_sym<4, 4>(c, a);
        else if (w.ri == 2 && w.warps == 4) launch_sym<2, 4>(c, a);
        else if (w.ri == 4 && w.warps == 2) launch_sym<4, 2>(c, a);
        else if (w.ri == 1 && w.warps == 4) launch_sym<1, 4>(c, a);
        else if (w.ri == 4 && w.warps == 8) launch_sym<4, 8>(c, a);
        else if (w.ri == 2 && w.warps == 8) launch_sym<2, 8>(c, a);
        else if (w.ri == 2 && w.warps == 1) launch_sym<2, 1>(c, a);
        else if (w.ri == 4 && w.warps == 1) launch_sym<4, 1>(c, a);
        else launch_sym<2, 2>(c, a);
    }
    if (!c->capturing) c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

} // namespace ljmd
