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
#include <cstring>
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
    long long *dbg;                  // optional per-warp {start, end, smid, items} (tuning)
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

// One 32-atom j-chunk against the RI*32 i-atoms of a warp.
//
// Rotation with RJ offsets per step: at step s (s = 0 .. 32/RJ - 1) lane l meets the RJ j-atoms
// (l + s + q*32/RJ) & 31, q = 0..RJ-1, so every step carries RI*RJ independent pair evaluations
// (the instruction-level parallelism a single warp needs: measured 252 / 355 / 490 cycles per
// step for 1 / 2 / 4 pairs, tools/chunk_cycles.py) and a chunk needs only 32/RJ serial steps.
// Each offset q has its own j-side accumulator set that rotates one lane per step (SHFL); at
// the end the RJ copies of every j-atom's force are gathered in the fixed order q = 0..RJ-1.
//
// The interaction body is branch-free: a pair outside the cutoff contributes exact zeros (its
// 1/r^2 is replaced by 0), so no per-pair branch serialises the dependency chains.  Unless
// DENSE, the body is skipped when no pair of the step is inside the cutoff.
template <int RI, int RJ, bool DIAG, bool DENSE, bool COUNT>
__device__ __forceinline__ void sym_chunk(const double (*sjb)[32], const double (&xi)[RI], const double (&yi)[RI],
                                          const double (&zi)[RI], double (&fxi)[RI], double (&fyi)[RI],
                                          double (&fzi)[RI], double &fjx_out, double &fjy_out, double &fjz_out,
                                          double &s12, double &s6, unsigned &ncut, double L, double invL, double rc2,
                                          int lane, int iw0, int j0, int n)
{
    constexpr int NS = 32 / RJ;
    double fjx[RJ], fjy[RJ], fjz[RJ];
#pragma unroll
    for (int q = 0; q < RJ; ++q) { fjx[q] = 0.0; fjy[q] = 0.0; fjz[q] = 0.0; }
#pragma unroll 2
    for (int s = 0; s < NS; ++s) {
        double dx[RJ][RI], dy[RJ][RI], dz[RJ][RI], r2[RJ][RI];
        bool in[RJ][RI];
        bool any = false;
#pragma unroll
        for (int q = 0; q < RJ; ++q) {
            const int jj = (lane + s + q * NS) & 31;
            const double xj = sjb[0][jj], yj = sjb[1][jj], zj = sjb[2][jj];
#pragma unroll
            for (int r = 0; r < RI; ++r) {
                dx[q][r] = mic(xi[r] - xj, L, invL); // lj_potential_energy.f90:119-126
                dy[q][r] = mic(yi[r] - yj, L, invL);
                dz[q][r] = mic(zi[r] - zj, L, invL);
                r2[q][r] = dist2(dx[q][r], dy[q][r], dz[q][r]); // :129, reference operation order
                in[q][r] = r2[q][r] < rc2;                      // :132
                if (DIAG) {
                    const int ig = iw0 + r * 32 + lane, jg = j0 + jj;
                    in[q][r] = in[q][r] && (jg > ig) && (jg < n);
                }
                any = any || in[q][r];
            }
        }
        if (DENSE || any) {
#pragma unroll
            for (int q = 0; q < RJ; ++q) {
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    // r2 = 0 can only occur on the diagonal (self pair); elsewhere r2 >= rc2 > 0
                    // is finite, its reciprocal is finite and is discarded by the select below.
                    double ir2 = fast_rcp((DIAG && !in[q][r]) ? 1.0 : r2[q][r]);
                    ir2 = in[q][r] ? ir2 : 0.0;
                    const double ir6 = __dmul_rn(__dmul_rn(ir2, ir2), ir2);
                    const double ir12 = __dmul_rn(ir6, ir6);
                    s12 = __dadd_rn(s12, ir12);
                    s6 = __dadd_rn(s6, ir6);
                    const double g = __dmul_rn(fma(2.0, ir12, -ir6), ir2); // -dU_r / r^2  (:143,148)
                    fxi[r] = fma(g, dx[q][r], fxi[r]); fyi[r] = fma(g, dy[q][r], fyi[r]); fzi[r] = fma(g, dz[q][r], fzi[r]);
                    fjx[q] = fma(-g, dx[q][r], fjx[q]); fjy[q] = fma(-g, dy[q][r], fjy[q]); fjz[q] = fma(-g, dz[q][r], fjz[q]);
                    if (COUNT) ncut += in[q][r] ? 1u : 0u;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RJ; ++q) {
            fjx[q] = __shfl_sync(0xffffffffu, fjx[q], (lane + 1) & 31);
            fjy[q] = __shfl_sync(0xffffffffu, fjy[q], (lane + 1) & 31);
            fjz[q] = __shfl_sync(0xffffffffu, fjz[q], (lane + 1) & 31);
        }
    }
    // copy q of j-atom m now sits in lane (m - (q+1)*NS) & 31; gather in the order q = 0..RJ-1
    if (RJ == 1) {
        fjx_out = fjx[0]; fjy_out = fjy[0]; fjz_out = fjz[0];
    } else {
        double tx = 0.0, ty = 0.0, tz = 0.0;
#pragma unroll
        for (int q = 0; q < RJ; ++q) {
            const int src = (lane - (q + 1) * NS) & 31;
            tx += __shfl_sync(0xffffffffu, fjx[q], src);
            ty += __shfl_sync(0xffffffffu, fjy[q], src);
            tz += __shfl_sync(0xffffffffu, fjz[q], src);
        }
        fjx_out = tx; fjy_out = ty; fjz_out = tz;
    }
}

template <int RI, int RJ, int W, bool DENSE, bool COUNT>
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
                sym_chunk<RI, RJ, false, DENSE, COUNT>(sj[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut,
                                            L, invL, rc2, lane, iw0, j0, n);
            else
                sym_chunk<RI, RJ, true, DENSE, COUNT>(sj[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut,
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
// "Independent warps" flavour of the symmetric kernel (rows of RI*32 atoms, W = 1 geometry):
// a CTA is just NW warps that never synchronise with each other.  Warp g of the launch takes the
// work items g, g + (number of warps), ... ; consecutive items go to different SMs.  Each warp
// stages its j-chunks in a private shared-memory slot (__syncwarp only) and writes the j-side
// partial of a chunk straight from registers.  Few fat CTAs instead of thousands of one-warp
// CTAs: no CTA-launch ramp, no barrier stalls.  Outputs are indexed by work item, so the result
// does not depend on which warp ran an item.
// ------------------------------------------------------------------------------------------
template <int RI, int RJ, bool DENSE, bool COUNT>
__global__ void __launch_bounds__(512) lj_pairs_symw_kernel(SymArgs a, int n_items)
{
    constexpr int TI = RI * 32;
    constexpr int MAXW = 16; // warps per CTA is a launch parameter (<= 16: 128 registers/thread)
    __shared__ double sj[MAXW][2][3][32];
    __shared__ double scratch[3 * MAXW];
    const int NW = blockDim.x >> 5;

    const int n = a.sc.n;
    const int rep = blockIdx.y;
    const double *rx = a.rx + (size_t)rep * n, *ry = a.ry + (size_t)rep * n,
                 *rz = a.rz + (size_t)rep * n;
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double (*sjw)[3][32] = sj[warp];

    double s12 = 0.0, s6 = 0.0;
    unsigned ncut = 0;
    const int nwarps_total = gridDim.x * NW;
    long long t_start = 0;
    int n_done = 0, n_diag = 0;
    if (a.dbg) t_start = clock64();
    for (int it = warp * gridDim.x + blockIdx.x; it < n_items; it += nwarps_total) {
        ++n_done;
        const int4 item = a.items[it];
        const int row = item.x, chunk0 = item.y, nchunks = item.z, split = item.w;
        const int iw0 = row * TI;
        double xi[RI], yi[RI], zi[RI], fxi[RI], fyi[RI], fzi[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int ig = iw0 + r * 32 + lane;
            const bool v = ig < n;
            xi[r] = v ? rx[ig] : 0.0; yi[r] = v ? ry[ig] : 0.0; zi[r] = v ? rz[ig] : 0.0;
            fxi[r] = 0.0; fyi[r] = 0.0; fzi[r] = 0.0;
        }
        {
            const int jg = chunk0 * 32 + lane;
            const bool v = jg < n;
            __syncwarp();
            sjw[0][0][lane] = v ? rx[jg] : 0.0;
            sjw[0][1][lane] = v ? ry[jg] : 0.0;
            sjw[0][2][lane] = v ? rz[jg] : 0.0;
            __syncwarp();
        }
        for (int k = 0; k < nchunks; ++k) {
            const int buf = k & 1;
            const int j0 = (chunk0 + k) * 32;
            double px = 0.0, py = 0.0, pz = 0.0; // next chunk, in flight during the compute
            if (k + 1 < nchunks) {
                const int jg = j0 + 32 + lane;
                if (jg < n) { px = rx[jg]; py = ry[jg]; pz = rz[jg]; }
            }
            double fjx = 0.0, fjy = 0.0, fjz = 0.0;
            const bool all = (j0 >= iw0 + TI) && (j0 + 32 <= n);
            if (!all) n_diag++;
            if (all)
                sym_chunk<RI, RJ, false, DENSE, COUNT>(sjw[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6,
                                                   ncut, L, invL, rc2, lane, iw0, j0, n);
            else
                sym_chunk<RI, RJ, true, DENSE, COUNT>(sjw[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6,
                                                  ncut, L, invL, rc2, lane, iw0, j0, n);
            // lane l holds the j-side force of atom j0 + l again
            const int jg = j0 + lane;
            if (jg < n) {
                double *fj = a.fJ + ((size_t)rep * a.n_rows + row) * 3 * n;
                fj[jg] = fjx; fj[(size_t)n + jg] = fjy; fj[2 * (size_t)n + jg] = fjz;
            }
            if (k + 1 < nchunks) {
                sjw[buf ^ 1][0][lane] = px; sjw[buf ^ 1][1][lane] = py; sjw[buf ^ 1][2][lane] = pz;
            }
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int ig = iw0 + r * 32 + lane;
            if (ig < n) {
                double *fp = a.fI + ((size_t)rep * a.max_splits + split) * 3 * n;
                fp[ig] = fxi[r]; fp[(size_t)n + ig] = fyi[r]; fp[2 * (size_t)n + ig] = fzi[r];
            }
        }
    }
    if (a.dbg && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *d = a.dbg + 4 * ((size_t)blockIdx.x * NW + warp);
        d[0] = t_start; d[1] = clock64(); d[2] = smid; d[3] = n_done + 256 * n_diag;
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
// Spatially sorted flavour of the independent-warp kernel (sparse cutoff, large N: the north star's
// "all-pairs at N = 32k").  Still every unordered pair i<j is distance-tested exactly once, like the
// reference loop (lj_potential_energy.f90:109-132) -- nothing is skipped -- but the resident atoms are
// kept in a spatially sorted order (k-d bisection into 32-atom leaves, done by the host at upload;
// any order gives the same set of pairs), and that makes two things cheap:
//   * minimum image (geometry_pbc.f90:86).  For a tile (128 i-atoms of a warp) x (32-atom j-chunk)
//     with bounding-box centres ci, cj and half widths hi, hj, let k = nint((ci-cj)/L) per component
//     and dx0 = (xi - xj) - L*k.  If |ci - cj - L*k| + hi + hj < L - rc, then for every pair of the
//     tile either |dx0| < L/2, where k IS the reference's dnint(dx/L) and dx0 is bit-identical to the
//     reference's dx - L*dnint(dx*invL) (L*k is exact, one rounding in the subtraction), or
//     |dx0| >= L/2 > rc while the true image is still farther than L - |dx0| > rc: the pair is
//     outside the cutoff for the reference and for dx0 alike.  So r2 < rc2 takes the reference's
//     decision for every pair and every in-cutoff pair has the reference's exact dx, dy, dz, r2 --
//     with ONE add per component (none when k = 0) instead of multiply + round + multiply-subtract.
//     Tiles that fail the test (never for a sorted homogeneous system) use the general path.
//   * the interaction stage.  A first pass distance-tests the 4096 pairs of a tile with the j-atom
//     broadcast from shared memory (no shuffles, 4 independent pairs per lane and j) and only records,
//     per 32 x 32 sub-tile, whether ANY pair is inside the cutoff (2.6 % of the sub-tiles at rho = 0.8,
//     rc = 2.5).  Only those sub-tiles run the rotation pass with the interaction body and the j-side
//     accumulators; its j-side partial is written, and flagged in `near`, only then.
// Work items are handed out dynamically (atomic counter); every output is indexed by work item or
// (row, chunk), so results do not depend on which warp ran what: bit-reproducible.
// ------------------------------------------------------------------------------------------
struct SortedArgs {
    SymArgs s;
    const double *cbox;  // [nchunks][6] centre xyz, half width xyz of every j-chunk
    unsigned char *near; // [nchunks][n_rows]
    unsigned *work;      // [2] next item, finished warps
    int nchunks;
    double lim;          // L - rc - margin
    long long rc2_bits;  // bit pattern of rc2: non-negative doubles order like their bit patterns
};

__global__ void __launch_bounds__(256) chunk_bbox_kernel(int n, const double *rx, const double *ry, const double *rz,
                                                         double *cbox)
{
    const int chunk = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (chunk >= (n + 31) / 32) return;
    const int j = chunk * 32 + lane;
    const bool v = j < n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double lo0 = v ? rx[j] : inf, lo1 = v ? ry[j] : inf, lo2 = v ? rz[j] : inf;
    double hi0 = v ? lo0 : -inf, hi1 = v ? lo1 : -inf, hi2 = v ? lo2 : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo0 = fmin(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmax(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
        lo1 = fmin(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmax(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
        lo2 = fmin(lo2, __shfl_xor_sync(0xffffffffu, lo2, o)); hi2 = fmax(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
    }
    if (lane == 0) {
        double *b = cbox + 6 * (size_t)chunk;
        b[0] = 0.5 * (lo0 + hi0); b[1] = 0.5 * (lo1 + hi1); b[2] = 0.5 * (lo2 + hi2);
        b[3] = 0.5 * (hi0 - lo0); b[4] = 0.5 * (hi1 - lo1); b[5] = 0.5 * (hi2 - lo2);
    }
}

// Pass 1 of a tile: distance test of all 128 x 32 pairs (lj_potential_energy.f90:119-132), j broadcast.
// SHIFT = false: no component of the tile has an image shift (42 % of the tiles), else one add per
// component.  Only two variants and a short loop body on purpose: the first version had one variant per
// shift pattern (8 x 200 instructions) and spent 37 % of its issue slots waiting for instruction fetch
// (ncu: stall_no_instruction, profiles/r2_ncu_cfg3_sorted_v1.txt).
// Returns the mask of 32 x 32 sub-tiles (i-atoms r*32 .. r*32+31 of the warp) that hold at least one
// pair inside the cutoff.
template <bool SHIFT>
__device__ __forceinline__ unsigned sorted_far_pass(const double (*sjb)[32], const double (&xi)[4], const double (&yi)[4],
                                                    const double (&zi)[4], double shx, double shy, double shz,
                                                    long long rc2b)
{
    bool hit[4] = {false, false, false, false};
#pragma unroll 2
    for (int jj = 0; jj < 32; ++jj) {
        const double xj = sjb[0][jj], yj = sjb[1][jj], zj = sjb[2][jj];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double dx = __dadd_rn(xi[r], -xj), dy = __dadd_rn(yi[r], -yj), dz = __dadd_rn(zi[r], -zj); // :119-121
            if (SHIFT) { // :124-126, see the header of this section
                dx = __dadd_rn(dx, shx);
                dy = __dadd_rn(dy, shy);
                dz = __dadd_rn(dz, shz);
            }
            const double r2 = dist2(dx, dy, dz);                  // :129
            hit[r] = hit[r] || (__double_as_longlong(r2) < rc2b); // :132 (r2 >= 0: integer compare, off the FP64 pipe)
        }
    }
    unsigned mask = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (__any_sync(0xffffffffu, hit[r])) mask |= 1u << r;
    return mask;
}

// Pass 2, only for the sub-tiles of `mask`: the rotation scheme of sym_chunk (lane l meets j-atom
// (l+s)&31, the j-side accumulators rotate through the warp) with the tile's image shift.
template <bool COUNT>
__device__ __forceinline__ void sorted_near_pass(unsigned mask, const double (*sjb)[32], const double (&xi)[4],
                                                 const double (&yi)[4], const double (&zi)[4], double (&fxi)[4],
                                                 double (&fyi)[4], double (&fzi)[4], double &fjx, double &fjy, double &fjz,
                                                 double &s12, double &s6, unsigned &ncut, double shx, double shy,
                                                 double shz, double rc2, int lane)
{
    fjx = 0.0; fjy = 0.0; fjz = 0.0;
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
        const int jj = (lane + s) & 31;
        const double xj = sjb[0][jj], yj = sjb[1][jj], zj = sjb[2][jj];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (mask & (1u << r)) { // warp-uniform
                const double dx = __dadd_rn(__dadd_rn(xi[r], -xj), shx);
                const double dy = __dadd_rn(__dadd_rn(yi[r], -yj), shy);
                const double dz = __dadd_rn(__dadd_rn(zi[r], -zj), shz);
                const double r2 = dist2(dx, dy, dz);
                const bool in = r2 < rc2;
                if (__any_sync(0xffffffffu, in)) {
                    double ir2 = fast_rcp(r2);
                    ir2 = in ? ir2 : 0.0;
                    const double ir6 = __dmul_rn(__dmul_rn(ir2, ir2), ir2);
                    const double ir12 = __dmul_rn(ir6, ir6);
                    s12 = __dadd_rn(s12, ir12);
                    s6 = __dadd_rn(s6, ir6);
                    const double g = __dmul_rn(fma(2.0, ir12, -ir6), ir2); // -dU_r / r^2  (:143,148)
                    fxi[r] = fma(g, dx, fxi[r]); fyi[r] = fma(g, dy, fyi[r]); fzi[r] = fma(g, dz, fzi[r]);
                    fjx = fma(-g, dx, fjx); fjy = fma(-g, dy, fjy); fjz = fma(-g, dz, fjz);
                    if (COUNT) ncut += in ? 1u : 0u;
                }
            }
        }
        fjx = __shfl_sync(0xffffffffu, fjx, (lane + 1) & 31);
        fjy = __shfl_sync(0xffffffffu, fjy, (lane + 1) & 31);
        fjz = __shfl_sync(0xffffffffu, fjz, (lane + 1) & 31);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(512) lj_pairs_sorted_kernel(SortedArgs A, int n_items)
{
    constexpr int RI = 4, TI = RI * 32, MAXW = 16;
    const SymArgs &a = A.s;
    __shared__ double sj[MAXW][2][3][32];
    const int NW = blockDim.x >> 5;
    const int n = a.sc.n;
    const double *rx = a.rx, *ry = a.ry, *rz = a.rz;
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double (*sjw)[3][32] = sj[warp];
    unsigned ncut = 0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    for (;;) {
        int it = 0;
        if (lane == 0) it = (int)atomicAdd(&A.work[0], 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const int4 item = a.items[it];
        const int row = item.x, chunk0 = item.y, nchunks = item.z, split = item.w;
        const int iw0 = row * TI;
        double xi[RI], yi[RI], zi[RI], fxi[RI], fyi[RI], fzi[RI];
        double lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int ig = iw0 + r * 32 + lane;
            const bool v = ig < n;
            xi[r] = v ? rx[ig] : 0.0; yi[r] = v ? ry[ig] : 0.0; zi[r] = v ? rz[ig] : 0.0;
            fxi[r] = 0.0; fyi[r] = 0.0; fzi[r] = 0.0;
            if (v) {
                lo0 = fmin(lo0, xi[r]); hi0 = fmax(hi0, xi[r]);
                lo1 = fmin(lo1, yi[r]); hi1 = fmax(hi1, yi[r]);
                lo2 = fmin(lo2, zi[r]); hi2 = fmax(hi2, zi[r]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo0 = fmin(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmax(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
            lo1 = fmin(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmax(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
            lo2 = fmin(lo2, __shfl_xor_sync(0xffffffffu, lo2, o)); hi2 = fmax(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
        }
        const double cix = 0.5 * (lo0 + hi0), ciy = 0.5 * (lo1 + hi1), ciz = 0.5 * (lo2 + hi2);
        const double hix = 0.5 * (hi0 - lo0), hiy = 0.5 * (hi1 - lo1), hiz = 0.5 * (hi2 - lo2);
        double s12 = 0.0, s6 = 0.0;
        {
            const int jg = chunk0 * 32 + lane;
            const bool v = jg < n;
            __syncwarp();
            sjw[0][0][lane] = v ? rx[jg] : 0.0;
            sjw[0][1][lane] = v ? ry[jg] : 0.0;
            sjw[0][2][lane] = v ? rz[jg] : 0.0;
            __syncwarp();
        }
        // centre / half width of the current chunk, one component per lane (lanes 0..5)
        double bcur = (lane < 6) ? A.cbox[6 * (size_t)chunk0 + lane] : 0.0;
        for (int k = 0; k < nchunks; ++k) {
            const int buf = k & 1;
            const int chunk = chunk0 + k, j0 = chunk * 32;
            double px = 0.0, py = 0.0, pz = 0.0, bnext = 0.0; // next chunk, in flight during the compute
            if (k + 1 < nchunks) {
                const int jg = j0 + 32 + lane;
                if (jg < n) { px = rx[jg]; py = ry[jg]; pz = rz[jg]; }
                if (lane < 6) bnext = A.cbox[6 * (size_t)(chunk + 1) + lane];
            }
            double fjx = 0.0, fjy = 0.0, fjz = 0.0;
            bool wrote = true;
            const bool all = (j0 >= iw0 + TI) && (j0 + 32 <= n);
            bool general = !all; // diagonal / ragged tile, or a tile that fails the shift test: general path
            double shx = 0.0, shy = 0.0, shz = 0.0;
            if (all) {
                // image shift of the tile and the test that makes it valid (header of this section)
                const double cjx = __shfl_sync(0xffffffffu, bcur, 0), cjy = __shfl_sync(0xffffffffu, bcur, 1),
                             cjz = __shfl_sync(0xffffffffu, bcur, 2), hjx = __shfl_sync(0xffffffffu, bcur, 3),
                             hjy = __shfl_sync(0xffffffffu, bcur, 4), hjz = __shfl_sync(0xffffffffu, bcur, 5);
                const double cx = cix - cjx, cy = ciy - cjy, cz = ciz - cjz;
                shx = -(L * rint(cx * invL)); shy = -(L * rint(cy * invL)); shz = -(L * rint(cz * invL));
                general = !((fabs(cx + shx) + (hix + hjx) < A.lim) && (fabs(cy + shy) + (hiy + hjy) < A.lim) &&
                            (fabs(cz + shz) + (hiz + hjz) < A.lim));
            }
            if (general) {
                // one instantiation for both cases: the per-pair masks of the diagonal variant (j > i, j < n)
                // are all true for a full off-diagonal tile
                sym_chunk<RI, 1, true, false, COUNT>(sjw[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut, L,
                                                     invL, rc2, lane, iw0, j0, n);
            } else {
                const bool noshift = (shx == 0.0) && (shy == 0.0) && (shz == 0.0);
                const unsigned mask = noshift ? sorted_far_pass<false>(sjw[buf], xi, yi, zi, shx, shy, shz, A.rc2_bits)
                                              : sorted_far_pass<true>(sjw[buf], xi, yi, zi, shx, shy, shz, A.rc2_bits);
                if (mask) sorted_near_pass<COUNT>(mask, sjw[buf], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut,
                                                  shx, shy, shz, rc2, lane);
                else wrote = false;
            }
            if (wrote) { // lane l holds the j-side force of atom j0 + l
                const int jg = j0 + lane;
                if (jg < n) {
                    double *fj = a.fJ + (size_t)row * 3 * n;
                    fj[jg] = fjx; fj[(size_t)n + jg] = fjy; fj[2 * (size_t)n + jg] = fjz;
                }
            }
            if (lane == 0) A.near[(size_t)chunk * a.n_rows + row] = wrote ? 1 : 0;
            if (k + 1 < nchunks) {
                sjw[buf ^ 1][0][lane] = px; sjw[buf ^ 1][1][lane] = py; sjw[buf ^ 1][2][lane] = pz;
            }
            bcur = bnext;
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int ig = iw0 + r * 32 + lane;
            if (ig < n) {
                double *fp = a.fI + (size_t)split * 3 * n;
                fp[ig] = fxi[r]; fp[(size_t)n + ig] = fyi[r]; fp[2 * (size_t)n + ig] = fzi[r];
            }
        }
        // scalar partials per WORK ITEM (not per warp: which warp ran the item varies run to run)
        s12 = warp_sum(s12);
        s6 = warp_sum(s6);
        if (lane == 0) {
            double *sp = a.spart + (size_t)it * 3;
            sp[0] = s12; sp[1] = s6; sp[2] = 0.0;
        }
    }
    if (COUNT) {
        const unsigned c = __reduce_add_sync(0xffffffffu, ncut);
        if (lane == 0 && a.pair_counts) atomicAdd(&a.pair_counts[1], (unsigned long long)c);
    }
    if (lane == 0) { // the last warp to run dry re-arms the counters for the next launch
        __threadfence();
        const unsigned d = atomicAdd(&A.work[1], 1u);
        if (d == gridDim.x * NW - 1) { A.work[0] = 0u; A.work[1] = 0u; }
    }
}

// ------------------------------------------------------------------------------------------
// Self-benchmark of the inner routine: `nw` warps per CTA, one CTA per SM, each warp runs
// `reps` chunks (32 rotation steps each) on synthetic in-cutoff data; reports the cycles per
// rotation step seen by warp 0 of CTA 0 (clock64).  Used to size RI / warps per SM.
// ------------------------------------------------------------------------------------------
template <int RI, int RJ, bool DENSE>
__global__ void sym_chunk_bench_kernel(double *out, long long *cycles, int reps, double L, double rc2)
{
    __shared__ double sj[32][3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double xi[RI], yi[RI], zi[RI], fxi[RI], fyi[RI], fzi[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
        xi[r] = 0.01 * lane + 0.3 * r; yi[r] = 0.02 * lane; zi[r] = 0.5 + 0.01 * r;
        fxi[r] = fyi[r] = fzi[r] = 0.0;
    }
    sj[warp][0][lane] = 1.5 + 0.011 * lane; sj[warp][1][lane] = 1.0 + 0.02 * lane; sj[warp][2][lane] = 0.7;
    __syncthreads();
    double fjx = 0, fjy = 0, fjz = 0, s12 = 0, s6 = 0;
    unsigned ncut = 0;
    const long long t0 = clock64();
    for (int k = 0; k < reps; ++k)
        sym_chunk<RI, RJ, false, DENSE, false>(sj[warp], xi, yi, zi, fxi, fyi, fzi, fjx, fjy, fjz, s12, s6, ncut, L, 1.0 / L,
                                           rc2, lane, 0, 1 << 20, 1 << 30);
    const long long t1 = clock64();
    double s = fjx + fjy + fjz + s12 + s6;
#pragma unroll
    for (int r = 0; r < RI; ++r) s += fxi[r] + fyi[r] + fzi[r];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int RI, int RJ>
static void bench_launch(int dense, int sms, int threads, double *d, long long *dc, int reps, double L, double rc2)
{
    if (dense) sym_chunk_bench_kernel<RI, RJ, true><<<sms, threads>>>(d, dc, reps, L, rc2);
    else sym_chunk_bench_kernel<RI, RJ, false><<<sms, threads>>>(d, dc, reps, L, rc2);
}

// cycles per 32x(32*ri) tile divided by 32: comparable across rj ("cycles per rotation of one lane")
int bench_chunk_cycles(int ri, int rj, int dense, int warps_per_cta, double *cycles_per_step)
{
    if (warps_per_cta < 1 || warps_per_cta > 32) return set_error(LJMD_ERR_ARG, "warps_per_cta out of range");
    double *d = nullptr;
    long long *dc = nullptr;
    LJMD_CUDA_CHECK(cudaMalloc(&d, 8));
    LJMD_CUDA_CHECK(cudaMalloc(&dc, 8));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int reps = 64, th = warps_per_cta * 32;
    const double L = 100.0, rc2 = dense ? 49.0 * 49.0 : 0.01;
    int ok = 1;
    for (int pass = 0; pass < 2 && ok; ++pass) {
        if (ri == 1 && rj == 1) bench_launch<1, 1>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 2 && rj == 1) bench_launch<2, 1>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 4 && rj == 1) bench_launch<4, 1>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 1 && rj == 2) bench_launch<1, 2>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 2 && rj == 2) bench_launch<2, 2>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 1 && rj == 4) bench_launch<1, 4>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 2 && rj == 4) bench_launch<2, 4>(dense, sms, th, d, dc, reps, L, rc2);
        else if (ri == 4 && rj == 2) bench_launch<4, 2>(dense, sms, th, d, dc, reps, L, rc2);
        else ok = 0;
        if (ok) LJMD_CUDA_CHECK(cudaDeviceSynchronize());
    }
    long long c = 0;
    if (ok) LJMD_CUDA_CHECK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(dc);
    if (!ok) return set_error(LJMD_ERR_ARG, "unsupported (ri, rj)");
    *cycles_per_step = (double)c / (reps * 32.0);
    return LJMD_OK;
}

// ------------------------------------------------------------------------------------------
// Planning
// ------------------------------------------------------------------------------------------
// Total number of (row, chunk) visits of the triangle.
static long long sym_total_chunk_visits(int n, int tcta)
{
    const int nchunks = (n + 31) / 32;
    const int nrows = (n + tcta - 1) / tcta;
    long long total = 0;
    for (int r = 0; r < nrows; ++r) total += nchunks - (long long)r * tcta / 32;
    return total;
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
        if (n >= 16384) { w.ri = 4; w.warps = 4; } else { w.ri = 2; w.warps = 1; }
        // fraction of all pairs inside the cutoff sphere: (4 pi / 3) (rc / L)^3
        const double frac = 4.18879020478639 * c->p.rc * c->p.rc * c->p.rc / c->p.volume;
        w.dense = frac > 0.15;
        if (const char *e = getenv("LJMD_SYM_DENSE")) w.dense = atoi(e) != 0;

        // tuning overrides (benchmark sweeps only)
        if (const char *e = getenv("LJMD_SYM_RI")) w.ri = atoi(e);
        if (const char *e = getenv("LJMD_SYM_W")) w.warps = atoi(e);
        w.rj = (n >= 16384) ? 1 : 2;
        if (const char *e = getenv("LJMD_SYM_RJ")) w.rj = atoi(e);
        // Sparse cutoff at large N: spatially sorted state + lj_pairs_sorted_kernel (rows of 128 atoms,
        // independent warps).  Needs one j-side partial plane per row: n_rows * 3 * n doubles.
        w.sorted = n >= 16384 && !w.dense && c->R == 1 && (c->o.flags & LJMD_FLAG_NO_SORT) == 0;
        if (const char *e = getenv("LJMD_SYM_SORTED")) w.sorted = w.sorted && atoi(e) != 0;
        if (w.sorted && sizeof(double) * 3.0 * n * ((n + 127) / 128) > 2.0 * 1073741824.0) w.sorted = false;
        if (w.sorted) { w.ri = 4; w.warps = 1; w.rj = 1; }
        // small systems: independent warps (rows of ri*32 atoms) packed into fat CTAs
        w.indep = (w.warps == 1);
        w.cta_warps = 16;
        if (const char *e = getenv("LJMD_SYM_INDEP")) w.indep = atoi(e) != 0;
        if (w.indep) w.warps = 1;
        const int tcta = w.ri * 32 * w.warps;
        const int nchunks = (n + 31) / 32;
        w.n_rows = (n + tcta - 1) / tcta;
        // Every row's chunk run is cut into near-equal items (lengths differ by at most one).
        //  * independent-warp kernel (small N): one item per warp, all warps resident at once
        //    (<= 16 warps per SM), so the item length is the smallest k with
        //    visits / k <= available warps; the kernel time is then k chunk-times for everybody.
        //  * CTA kernel (large N): about 16 chunks per item, many waves.
        const long long total = sym_total_chunk_visits(n, tcta);
        long long target;
        if (w.indep) {
            // All 16 warps of every SM stay resident and each takes >= 2 items when there is
            // enough work (the items a warp runs back to back balance the tail); small systems
            // get single-chunk items.  Measured per-warp timelines: tools/warp_times.py.
            const long long slots = std::max<long long>(1, (long long)sms * 16 / std::max(1, c->R));
            const long long kk = std::max<long long>(1, total / (2 * slots));
            target = (total + kk - 1) / kk;
        } else {
            target = (total + 15) / 16;
            const long long min_items = (long long)sms * 4 / std::max(1, c->R);
            if (target < min_items) target = min_items;
        }
        if (const char *e = getenv("LJMD_SYM_ITEMS")) target = atoll(e);
        if (target < 1) target = 1;
        std::vector<int4> items;
        std::vector<int> row_splits(w.n_rows, 0);
        int max_splits = 0;
        for (int r = 0; r < w.n_rows; ++r) {
            const int first = r * tcta / 32;
            const int m = nchunks - first;
            long long kr = (m * target + total / 2) / total;
            if (kr < 1) kr = 1;
            if (kr > m) kr = m;
            int c0 = first;
            if (w.sorted) {
                // 32-chunk items, then the rest of the row in pieces of <= 8 chunks: the items are handed
                // out dynamically, longest first, so the short ones level the tail
                // Pieces of at most 8 chunks, 4 next to the diagonal (the diagonal tiles take the general path and
                // the spatially sorted order puts most of a row's interacting sub-tiles there), and the far end of
                // the row in ever smaller pieces.  The items are handed out dynamically in order of estimated cost.
                // Sizing rule (measured, profiles/r2_ncu_cfg3_sorted_v2/v3.txt): 2368 resident warps share 125 k
                // tiles at N = 32k, 53 per warp; with 32-chunk pieces the SMs ran dry between 79 % and 100 % of
                // the kernel time -- the largest piece has to be a small fraction of a warp's share.
                static const int tail_pieces[] = {4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1};
                int left = m, s = 0;
                for (int q = 0; q < 2 && left > 24; ++q) {
                    const int len = std::min(left - 24, 4);
                    items.push_back(make_int4(r, c0, len, s++));
                    c0 += len; left -= len;
                }
                while (left > 24) {
                    const int len = std::min(left - 24, 8);
                    items.push_back(make_int4(r, c0, len, s++));
                    c0 += len; left -= len;
                }
                for (int q = 0; q < 11 && left > 0; ++q) {
                    const int len = std::min(left, tail_pieces[q]);
                    items.push_back(make_int4(r, c0, len, s++));
                    c0 += len; left -= len;
                }
                if (left > 0) { items.push_back(make_int4(r, c0, left, s++)); c0 += left; left = 0; }
                kr = 0; // nothing left for the generic splitter below
                row_splits[r] = s;
                max_splits = std::max(max_splits, s);
            }
            for (int s = 0; s < (int)kr; ++s) {
                const int len = m / (int)kr + (s < m % (int)kr ? 1 : 0);
                int4 it;
                it.x = r; it.y = c0; it.z = len; it.w = s;
                items.push_back(it);
                c0 += len;
            }
            if (!w.sorted) {
                row_splits[r] = (int)kr;
                max_splits = std::max(max_splits, (int)kr);
            }
        }
        w.chunks_per_item = 0; // unused: per-row split counts live in w.row_splits
        LJMD_CUDA_CHECK(cudaMalloc(&w.row_splits, sizeof(int) * w.n_rows));
        LJMD_CUDA_CHECK(cudaMemcpy(w.row_splits, row_splits.data(), sizeof(int) * w.n_rows, cudaMemcpyHostToDevice));
        // heavy items first: full-length runs before the short tails (stable => deterministic)
        if (w.sorted) {
            // estimated cost in far-tile units: chunks at offset d from the row's diagonal weigh 5 (d < 4: general
            // path), 2 (d < 32), 1.5 (d < 128), 1 (beyond) -- the measured density of interacting sub-tiles
            const int cpr = tcta / 32; // chunks per row of i-atoms
            auto cost = [cpr](const int4 &it) {
                double c = 0.0;
                for (int k = 0; k < it.z; ++k) {
                    const int d = it.y + k - it.x * cpr;
                    c += d < 4 ? 5.0 : d < 32 ? 2.0 : d < 128 ? 1.5 : 1.0;
                }
                return c;
            };
            std::stable_sort(items.begin(), items.end(), [&](const int4 &x, const int4 &y) { return cost(x) > cost(y); });
        } else {
            std::stable_sort(items.begin(), items.end(), [](const int4 &x, const int4 &y) { return x.z > y.z; });
        }
        w.n_items = (int)items.size();
        w.n_splits = max_splits;
        w.n_ctas = w.n_items;
        if (w.sorted) {
            w.n_ctas = sms;
            w.cta_warps = 16;
            const int nch = (n + 31) / 32;
            LJMD_CUDA_CHECK(cudaMalloc(&w.cbox, sizeof(double) * 6 * (size_t)nch));
            LJMD_CUDA_CHECK(cudaMalloc(&w.near, (size_t)w.n_rows * nch));
            LJMD_CUDA_CHECK(cudaMemsetAsync(w.near, 0, (size_t)w.n_rows * nch, c->stream));
            LJMD_CUDA_CHECK(cudaMalloc(&w.work, sizeof(unsigned) * 2));
            LJMD_CUDA_CHECK(cudaMemsetAsync(w.work, 0, sizeof(unsigned) * 2, c->stream));
        } else if (w.indep) {
            // fat CTAs of 16 independent warps, one CTA per SM when there is enough work
            int ctas = (w.n_items + 15) / 16;
            if (c->R == 1) ctas = std::min(ctas, sms);
            if (const char *e = getenv("LJMD_SYM_CTAS")) ctas = atoi(e);
            ctas = std::max(1, ctas);
            w.cta_warps = std::min(16, (w.n_items + ctas - 1) / ctas);
            if (const char *e = getenv("LJMD_SYM_NW")) w.cta_warps = std::max(1, std::min(16, atoi(e)));
            w.n_ctas = ctas;
        }
        LJMD_CUDA_CHECK(cudaMalloc(&w.items, sizeof(int4) * items.size()));
        LJMD_CUDA_CHECK(cudaMemcpy(w.items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice));
        const size_t ibytes = sizeof(double) * (size_t)c->R * max_splits * 3 * n;
        const size_t jbytes = sizeof(double) * (size_t)c->R * w.n_rows * 3 * n;
        LJMD_CUDA_CHECK(cudaMalloc(&w.fpart, ibytes));
        LJMD_CUDA_CHECK(cudaMalloc(&w.fJ, jbytes));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.fpart, 0, ibytes, c->stream));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.fJ, 0, jbytes, c->stream));
    }
    if (w.variant == 1 && w.indep && getenv("LJMD_SYM_DEBUG")) {
        LJMD_CUDA_CHECK(cudaMalloc(&w.dbg, sizeof(long long) * 4 * (size_t)w.n_ctas * w.cta_warps));
        LJMD_CUDA_CHECK(cudaMemset(w.dbg, 0, sizeof(long long) * 4 * (size_t)w.n_ctas * w.cta_warps));
    }
    w.n_spart = (w.variant == 1 && w.indep && !w.sorted) ? w.n_ctas : w.n_items;
    const size_t sbytes = sizeof(double) * (size_t)c->R * w.n_spart * 3;
    LJMD_CUDA_CHECK(cudaMalloc(&w.spart, sbytes));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.spart, 0, sbytes, c->stream));
    return LJMD_OK;
}

template <int RI, int RJ>
static void launch_symw(ljmd_ctx *c, const SymArgs &a)
{
    dim3 grid(c->pw.n_ctas, c->R);
    const int ni = c->pw.n_items, th = c->pw.cta_warps * 32;
    if (a.pair_counts) lj_pairs_symw_kernel<RI, RJ, false, true><<<grid, th, 0, c->stream>>>(a, ni);
    else if (c->pw.dense) lj_pairs_symw_kernel<RI, RJ, true, false><<<grid, th, 0, c->stream>>>(a, ni);
    else lj_pairs_symw_kernel<RI, RJ, false, false><<<grid, th, 0, c->stream>>>(a, ni);
}

template <int RI, int RJ, int W>
static void launch_sym(ljmd_ctx *c, const SymArgs &a)
{
    dim3 grid(c->pw.n_items, c->R);
    if (a.pair_counts) lj_pairs_sym_kernel<RI, RJ, W, false, true><<<grid, W * 32, 0, c->stream>>>(a);
    else if (c->pw.dense) lj_pairs_sym_kernel<RI, RJ, W, true, false><<<grid, W * 32, 0, c->stream>>>(a);
    else lj_pairs_sym_kernel<RI, RJ, W, false, false><<<grid, W * 32, 0, c->stream>>>(a);
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
        a.dbg = w.dbg;
        const int key = w.sorted ? -1 : (w.indep ? 10000 : 0) + w.ri * 1000 + w.rj * 100 + (w.indep ? 0 : w.warps);
        switch (key) {
        case -1: {
            SortedArgs A;
            A.s = a;
            A.cbox = w.cbox; A.near = w.near; A.work = w.work;
            A.nchunks = (c->sc.n + 31) / 32;
            A.lim = (c->sc.L - c->p.rc) * (1.0 - 1e-9);
            const double rc2 = c->sc.rc2;
            memcpy(&A.rc2_bits, &rc2, sizeof(double));
            chunk_bbox_kernel<<<(A.nchunks * 32 + 255) / 256, 256, 0, c->stream>>>(c->sc.n, px, py, pz, w.cbox);
            if (a.pair_counts) lj_pairs_sorted_kernel<true><<<w.n_ctas, 512, 0, c->stream>>>(A, w.n_items);
            else lj_pairs_sorted_kernel<false><<<w.n_ctas, 512, 0, c->stream>>>(A, w.n_items);
            if (!c->capturing) c->launches++;
            break;
        }
        case 12100: launch_symw<2, 1>(c, a); break;
        case 12200: launch_symw<2, 2>(c, a); break;
        case 11400: launch_symw<1, 4>(c, a); break;
        case 11200: launch_symw<1, 2>(c, a); break;
        case 14100: launch_symw<4, 1>(c, a); break;
        case 4104: launch_sym<4, 1, 4>(c, a); break;
        case 4204: launch_sym<4, 2, 4>(c, a); break;
        case 2204: launch_sym<2, 2, 4>(c, a); break;
        case 2104: launch_sym<2, 1, 4>(c, a); break;
        case 2208: launch_sym<2, 2, 8>(c, a); break;
        case 4108: launch_sym<4, 1, 8>(c, a); break;
        default:
            return set_error(LJMD_ERR_ARG, "unsupported symmetric-kernel geometry RI=%d RJ=%d W=%d indep=%d", w.ri, w.rj,
                             w.indep ? w.cta_warps : w.warps, (int)w.indep);
        }
    }
    if (!c->capturing) c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

} // namespace ljmd
