// md_analysis.cu -- trajectory consumers: the pair-distance
// histogram behind g(r) and the time-origin averaged MSD(tau) / VACF(tau), i.e. the O(n^2) and
// O(n_snap^2 n) loops of /root/reference/scripts/md_one_run_analysis.py:404-489,537-598.
//
// Arithmetic contract (compiled with -fmad=false; every operation below is a separate IEEE
// rounding, like the numpy expressions it mirrors):
//   RDF  (:572-584)  d = x_j - x_i;  d = d - L*rint(d/L)  (division, not a multiply by 1/L; rint
//        = ties-to-even like np.rint);  r = sqrt((dx*dx + dy*dy) + dz*dz);  counted when r < rmax in
//        bin (int)(r/dr).  Integer counts => the histogram is BIT-EXACT and order independent.
//   MSD  (:432-438)  mean over particles of (dx*dx + dy*dy) + dz*dz, d = ru(t0+lag) - ru(t0); the
//        means are added over the time origins t0 = 0, stride, 2 stride, ... in that order and
//        divided by the number of origins.  VACF (:472-484) likewise with
//        (vx*v0x + vy*v0y) + vz*v0z.  numpy's mean uses pairwise summation, the kernel a fixed
//        lane-strided order: equal to ~1e-15 relative, not bitwise (tests: 1e-12).
//
// Layout: trajectories are [n_snap][n] row-major per component, like the arrays the reference's
// read_rva returns (:355-390).
#include <stdint.h>
#include <stdlib.h>

#include "ljmd_internal.h"
#include "reduce.cuh"

namespace ljmd {

// ------------------------------------------------------------------------------------------
// RDF pair histogram
// ------------------------------------------------------------------------------------------
constexpr int RDF_T = 128;         // atoms per tile side
constexpr int RDF_BS = 256;        // threads per CTA: lane-distinct j, 64 i-atoms per thread
constexpr int RDF_SMEM_BINS = 8192; // histogram kept in shared memory up to this many bins

struct RdfArgs {
    const double *x, *y, *z; // [n_snap][n]
    int n, n_snap, nbins, ntile;
    double L, rmax, dr;
    unsigned long long *hist; // [nbins] unordered pairs
    unsigned int *bad;        // pairs with r < rmax whose bin index came out >= nbins
};

// One CTA = one pair of atom tiles (ti <= tj) of one snapshot (blockIdx.y strides the snapshots).
template <bool SMEM_HIST>
__global__ void __launch_bounds__(RDF_BS) rdf_hist_kernel(RdfArgs a)
{
    extern __shared__ unsigned sh_hist[];
    __shared__ double si[3][RDF_T], sj[3][RDF_T];

    int p = blockIdx.x, ti = 0, rowlen = a.ntile; // upper-triangle decode (ntile is small)
    while (p >= rowlen) { p -= rowlen; ++ti; --rowlen; }
    const int tj = ti + p;
    const int i0 = ti * RDF_T, j0 = tj * RDF_T;
    const int jl = threadIdx.x & (RDF_T - 1), ih = (threadIdx.x / RDF_T) * (RDF_T / 2);
    const int gj = j0 + jl;

    if (SMEM_HIST)
        for (int b = threadIdx.x; b < a.nbins; b += RDF_BS) sh_hist[b] = 0u;

    for (int s = blockIdx.y; s < a.n_snap; s += gridDim.y) {
        const size_t row = (size_t)s * a.n;
        __syncthreads(); // previous snapshot's tiles consumed; histogram zeroed
        if (threadIdx.x < RDF_T) {
            const int g = i0 + threadIdx.x;
            const bool v = g < a.n;
            si[0][threadIdx.x] = v ? a.x[row + g] : 0.0;
            si[1][threadIdx.x] = v ? a.y[row + g] : 0.0;
            si[2][threadIdx.x] = v ? a.z[row + g] : 0.0;
        } else {
            const int l = threadIdx.x - RDF_T, g = j0 + l;
            const bool v = g < a.n;
            sj[0][l] = v ? a.x[row + g] : 0.0;
            sj[1][l] = v ? a.y[row + g] : 0.0;
            sj[2][l] = v ? a.z[row + g] : 0.0;
        }
        __syncthreads();
        const double xj = sj[0][jl], yj = sj[1][jl], zj = sj[2][jl];
        if (gj < a.n) {
#pragma unroll 4
            for (int k = 0; k < RDF_T / 2; ++k) {
                const int il = ih + k, gi = i0 + il;
                if (gi >= a.n || gj <= gi) continue; // unique pairs i < j (:571-574)
                double dx = __dsub_rn(xj, si[0][il]);
                double dy = __dsub_rn(yj, si[1][il]);
                double dz = __dsub_rn(zj, si[2][il]);
                dx = __dsub_rn(dx, __dmul_rn(a.L, rint(__ddiv_rn(dx, a.L)))); // :575-577
                dy = __dsub_rn(dy, __dmul_rn(a.L, rint(__ddiv_rn(dy, a.L))));
                dz = __dsub_rn(dz, __dmul_rn(a.L, rint(__ddiv_rn(dz, a.L))));
                const double r = __dsqrt_rn(
                    __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))); // :578
                if (r < a.rmax) {                                   // :579
                    const long long b = (long long)__ddiv_rn(r, a.dr); // :580 astype(int): truncation
                    if (b < a.nbins) {
                        if (SMEM_HIST) atomicAdd(&sh_hist[(int)b], 1u);
                        else atomicAdd(&a.hist[b], 1ull);
                    } else {
                        atomicAdd(a.bad, 1u); // the reference's np.add.at raises IndexError here
                    }
                }
            }
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int b = threadIdx.x; b < a.nbins; b += RDF_BS) {
            const unsigned c = sh_hist[b];
            if (c) atomicAdd(&a.hist[b], (unsigned long long)c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// MSD / VACF: one warp per lag, origins in increasing order, no barriers
// ------------------------------------------------------------------------------------------
constexpr int CORR_BS = 256;

struct CorrArgs {
    const double *ax, *ay, *az; // [n_snap][n]: unwrapped positions (MSD) or velocities (VACF)
    int n, n_snap, max_lag, stride, kind; // kind 0 = MSD, 1 = VACF
    double *out;                          // [max_lag + 1]
};

template <int KIND>
__device__ __forceinline__ double corr_row_mean(const CorrArgs &a, int t0, int t1, int lane)
{
    const double *x0 = a.ax + (size_t)t0 * a.n, *y0 = a.ay + (size_t)t0 * a.n, *z0 = a.az + (size_t)t0 * a.n;
    const double *x1 = a.ax + (size_t)t1 * a.n, *y1 = a.ay + (size_t)t1 * a.n, *z1 = a.az + (size_t)t1 * a.n;
    double s = 0.0;
    for (int p = lane; p < a.n; p += 32) {
        double t;
        if (KIND == 0) { // :432-435
            const double dx = __dsub_rn(x1[p], x0[p]), dy = __dsub_rn(y1[p], y0[p]), dz = __dsub_rn(z1[p], z0[p]);
            t = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        } else {         // :477-482
            t = __dadd_rn(__dadd_rn(__dmul_rn(x1[p], x0[p]), __dmul_rn(y1[p], y0[p])), __dmul_rn(z1[p], z0[p]));
        }
        s = __dadd_rn(s, t);
    }
    return __ddiv_rn(warp_sum(s), (double)a.n); // np.mean(..., axis=1)
}

template <int KIND>
__global__ void __launch_bounds__(CORR_BS) time_corr_kernel(CorrArgs a)
{
    const int lane = threadIdx.x & 31;
    const int lag = blockIdx.x * (CORR_BS / 32) + (threadIdx.x >> 5);
    if (lag > a.max_lag) return; // whole warp
    double acc = 0.0;
    long long cnt = 0;
    if (a.n_snap < 2) {
        // :416-417 MSD of a single snapshot is [0]; :456-458 VACF is [<v0.v0>]
        if (KIND == 1) { acc = corr_row_mean<1>(a, 0, 0, lane); cnt = 1; }
    } else {
        for (int t0 = 0; t0 < a.n_snap - 1; t0 += a.stride) {  // :428 / :469
            const int lmax = min(a.max_lag, (a.n_snap - 1) - t0); // :429
            if (lmax <= 0) continue;                              // :430-431
            if (lag > lmax) break;                                // lmax never grows again
            acc = __dadd_rn(acc, corr_row_mean<KIND>(a, t0, t0 + lag, lane)); // :437 msd[:L+1] += d2
            ++cnt;                                                            // :438
        }
    }
    if (lane == 0) a.out[lag] = cnt > 0 ? __ddiv_rn(acc, (double)cnt) : 0.0; // :440-441
}

// ------------------------------------------------------------------------------------------
// Opt-in second version (LJMD_CORR_TILED=1; written after the round's GPU budget was spent, never
// run yet): register-tiled.  A warp owns CT_TO origins x CT_TL lags, lanes span the atoms, so a
// row element is loaded once per warp for up to CT_TL (origin row) or several (target row,
// stride 1: rows o + l coincide) of the 32 pairs instead of once per pair: ~4x less L1/L2
// traffic than one warp per lag, and 32 independent accumulation chains per lane.
// Each (origin, lag) mean is computed with exactly the arithmetic and order of corr_row_mean
// and written to dmat[origin][lag]; time_corr_reduce_kernel then adds the origins in increasing
// order per lag => results are bit-identical to time_corr_kernel.
// ------------------------------------------------------------------------------------------
constexpr int CT_TO = 4, CT_TL = 8; // per-warp tile: origins x lags
constexpr int CT_WO = 2, CT_WL = 4; // warps of a CTA along origins / lags (8 warps)

template <int KIND, bool STRIDE1>
__global__ void __launch_bounds__(CORR_BS, 2) time_corr_tiled_kernel(CorrArgs a, double *dmat, int n_orig)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = (blockIdx.y * CT_WO + warp / CT_WL) * CT_TO; // first origin index of this warp
    const int l0 = (blockIdx.x * CT_WL + warp % CT_WL) * CT_TL; // first lag of this warp
    if (o0 >= n_orig || l0 > a.max_lag) return;                 // whole warp
    const int stride = STRIDE1 ? 1 : a.stride;
    const int last = a.n_snap - 1;
    const size_t n = (size_t)a.n;
    // Rows are clamped into the trajectory so that every load is in bounds; pairs that do not
    // exist (origin index >= n_orig, lag beyond the origin's range) are dropped when storing.
    const int base = o0 * stride + l0;
    double acc[CT_TO][CT_TL];
#pragma unroll
    for (int o = 0; o < CT_TO; ++o)
#pragma unroll
        for (int l = 0; l < CT_TL; ++l) acc[o][l] = 0.0;

    for (int p = lane; p < a.n; p += 32) {
        double x0[CT_TO], y0[CT_TO], z0[CT_TO];
#pragma unroll
        for (int o = 0; o < CT_TO; ++o) {
            const size_t r = (size_t)min((o0 + o) * stride, last) * n + p;
            x0[o] = a.ax[r]; y0[o] = a.ay[r]; z0[o] = a.az[r];
        }
#pragma unroll
        for (int o = 0; o < CT_TO; ++o) {
#pragma unroll
            for (int l = 0; l < CT_TL; ++l) {
                // stride 1: the row depends on o + l only, so the unrolled loads of equal o + l merge
                const int row = STRIDE1 ? min(base + (o + l), last) : min(base + o * stride + l, last);
                const size_t r = (size_t)row * n + p;
                const double x1 = a.ax[r], y1 = a.ay[r], z1 = a.az[r];
                double t;
                if (KIND == 0) {
                    const double dx = __dsub_rn(x1, x0[o]), dy = __dsub_rn(y1, y0[o]), dz = __dsub_rn(z1, z0[o]);
                    t = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                } else {
                    t = __dadd_rn(__dadd_rn(__dmul_rn(x1, x0[o]), __dmul_rn(y1, y0[o])), __dmul_rn(z1, z0[o]));
                }
                acc[o][l] = __dadd_rn(acc[o][l], t);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < CT_TO; ++o) {
        const int oi = o0 + o;
        const int lmax = min(a.max_lag, last - oi * stride); // :429
#pragma unroll
        for (int l = 0; l < CT_TL; ++l) {
            const double m = __ddiv_rn(warp_sum(acc[o][l]), (double)a.n);
            if (lane == 0 && oi < n_orig && lmax > 0 && l0 + l <= lmax)
                dmat[(size_t)oi * (a.max_lag + 1) + (l0 + l)] = m;
        }
    }
}

// out[lag] = mean over the origins that reach `lag`, added in increasing origin order (:437-441).
__global__ void __launch_bounds__(CORR_BS) time_corr_reduce_kernel(CorrArgs a, const double *dmat, int n_orig)
{
    const int lag = blockIdx.x * CORR_BS + threadIdx.x;
    if (lag > a.max_lag) return;
    double acc = 0.0;
    long long cnt = 0;
    for (int o = 0; o < n_orig; ++o) {
        const int lmax = min(a.max_lag, (a.n_snap - 1) - o * a.stride);
        if (lmax <= 0) continue;
        if (lag > lmax) break;
        acc = __dadd_rn(acc, dmat[(size_t)o * (a.max_lag + 1) + lag]);
        ++cnt;
    }
    a.out[lag] = cnt > 0 ? __ddiv_rn(acc, (double)cnt) : 0.0;
}

// Device buffers released on every return path.
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

struct AnalysisDevice {
    int prev = -1;
    bool switched = false;
    int enter(int dev)
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return set_error(LJMD_ERR_CUDA, "no CUDA device available (this library has no CPU fallback).");
        }
        LJMD_CUDA_CHECK(cudaGetDevice(&prev));
        if (dev < 0) dev = prev;
        if (dev >= ndev) return set_error(LJMD_ERR_ARG, "device %d out of range.", dev);
        if (dev != prev) { LJMD_CUDA_CHECK(cudaSetDevice(dev)); switched = true; }
        return LJMD_OK;
    }
    ~AnalysisDevice() { if (switched) cudaSetDevice(prev); }
};

static int upload3(DevBuf &b, const double *x, const double *y, const double *z, size_t count)
{
    LJMD_CUDA_CHECK(cudaMalloc(&b.p, sizeof(double) * 3 * count));
    double *d = b.as<double>();
    LJMD_CUDA_CHECK(cudaMemcpy(d, x, sizeof(double) * count, cudaMemcpyHostToDevice));
    LJMD_CUDA_CHECK(cudaMemcpy(d + count, y, sizeof(double) * count, cudaMemcpyHostToDevice));
    LJMD_CUDA_CHECK(cudaMemcpy(d + 2 * count, z, sizeof(double) * count, cudaMemcpyHostToDevice));
    return LJMD_OK;
}

struct EventPair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

} // namespace ljmd

using namespace ljmd;

extern "C" int ljmd_rdf_histogram(const double *rx, const double *ry, const double *rz, int32_t n_snap, int32_t n,
                                  double box_length, int32_t nbins, double rmax, int64_t *hist, int32_t device,
                                  double *kernel_ms)
{
    clear_error();
    if (!rx || !ry || !rz || !hist) return set_error(LJMD_ERR_NULL, "ljmd_rdf_histogram(): NULL array.");
    if (n_snap < 1 || n < 2) return set_error(LJMD_ERR_ARG, "ljmd_rdf_histogram(): need n_snap >= 1 and n >= 2.");
    if (!(box_length > 0.0)) return set_error(LJMD_ERR_BOX, "ljmd_rdf_histogram(): box_length must be > 0.");
    if (nbins < 1 || !(rmax > 0.0)) return set_error(LJMD_ERR_ARG, "ljmd_rdf_histogram(): need nbins >= 1 and rmax > 0.");
    AnalysisDevice dev;
    LJMD_TRY(dev.enter(device));

    const size_t count = (size_t)n_snap * n;
    DevBuf pos, dh, dbad;
    LJMD_TRY(upload3(pos, rx, ry, rz, count));
    LJMD_CUDA_CHECK(cudaMalloc(&dh.p, sizeof(unsigned long long) * nbins));
    LJMD_CUDA_CHECK(cudaMemset(dh.p, 0, sizeof(unsigned long long) * nbins));
    LJMD_CUDA_CHECK(cudaMalloc(&dbad.p, sizeof(unsigned)));
    LJMD_CUDA_CHECK(cudaMemset(dbad.p, 0, sizeof(unsigned)));

    RdfArgs a;
    a.x = pos.as<double>(); a.y = a.x + count; a.z = a.y + count;
    a.n = n; a.n_snap = n_snap; a.nbins = nbins;
    a.ntile = (n + RDF_T - 1) / RDF_T;
    a.L = box_length; a.rmax = rmax;
    a.dr = rmax / (double)nbins; // :552
    a.hist = dh.as<unsigned long long>();
    a.bad = dbad.as<unsigned>();
    const long long ntp = (long long)a.ntile * (a.ntile + 1) / 2;
    if (ntp > 2147483647LL) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_rdf_histogram(): too many atoms (%d).", n);
    const dim3 grid((unsigned)ntp, (unsigned)(n_snap < 65535 ? n_snap : 65535));
    EventPair ev;
    LJMD_CUDA_CHECK(cudaEventCreate(&ev.e0));
    LJMD_CUDA_CHECK(cudaEventCreate(&ev.e1));
    LJMD_CUDA_CHECK(cudaEventRecord(ev.e0, 0));
    if (nbins <= RDF_SMEM_BINS)
        rdf_hist_kernel<true><<<grid, RDF_BS, sizeof(unsigned) * nbins>>>(a);
    else
        rdf_hist_kernel<false><<<grid, RDF_BS>>>(a);
    LJMD_CUDA_CHECK(cudaGetLastError());
    LJMD_CUDA_CHECK(cudaEventRecord(ev.e1, 0));
    LJMD_CUDA_CHECK(cudaDeviceSynchronize());
    if (kernel_ms) {
        float ms = 0.f;
        LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        *kernel_ms = ms;
    }
    unsigned bad = 0;
    LJMD_CUDA_CHECK(cudaMemcpy(&bad, dbad.p, sizeof(unsigned), cudaMemcpyDeviceToHost));
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "histogram element size");
    LJMD_CUDA_CHECK(cudaMemcpy(hist, dh.p, sizeof(int64_t) * nbins, cudaMemcpyDeviceToHost));
    if (bad)
        return set_error(LJMD_ERR_ARG,
                         "ljmd_rdf_histogram(): %u pair(s) with r < rmax fall into bin index nbins "
                         "(r/dr rounds up); the reference raises IndexError for this input.", bad);
    return LJMD_OK;
}

extern "C" int ljmd_time_correlation(const double *ax, const double *ay, const double *az, int32_t n_snap, int32_t n,
                                     int32_t max_lag, int32_t origin_stride, int32_t kind, double *out,
                                     int32_t device, double *kernel_ms)
{
    clear_error();
    if (!ax || !ay || !az || !out) return set_error(LJMD_ERR_NULL, "ljmd_time_correlation(): NULL array.");
    if (n_snap < 1 || n < 1) return set_error(LJMD_ERR_ARG, "ljmd_time_correlation(): need n_snap >= 1 and n >= 1.");
    if (kind != LJMD_CORR_MSD && kind != LJMD_CORR_VACF)
        return set_error(LJMD_ERR_ARG, "ljmd_time_correlation(): kind must be LJMD_CORR_MSD or LJMD_CORR_VACF.");
    // :420-423 / :461-464
    if (max_lag < 0 || max_lag > n_snap - 1) max_lag = n_snap - 1;
    if (n_snap < 2) max_lag = 0;
    if (origin_stride < 1) origin_stride = 1;
    if (origin_stride > n_snap) origin_stride = n_snap; // only t0 = 0 either way; keeps index products small
    AnalysisDevice dev;
    LJMD_TRY(dev.enter(device));

    const size_t count = (size_t)n_snap * n;
    DevBuf traj, dout;
    LJMD_TRY(upload3(traj, ax, ay, az, count));
    LJMD_CUDA_CHECK(cudaMalloc(&dout.p, sizeof(double) * (max_lag + 1)));
    CorrArgs a;
    a.ax = traj.as<double>(); a.ay = a.ax + count; a.az = a.ay + count;
    a.n = n; a.n_snap = n_snap; a.max_lag = max_lag; a.stride = origin_stride; a.kind = kind;
    a.out = dout.as<double>();
    const int lags_per_cta = CORR_BS / 32;
    const unsigned grid = (unsigned)((max_lag + 1 + lags_per_cta - 1) / lags_per_cta);
    const char *tiled_env = getenv("LJMD_CORR_TILED");
    const int n_orig = (n_snap - 1 + origin_stride - 1) / origin_stride; // len(range(0, n_snap-1, stride))
    const size_t dmat_bytes = sizeof(double) * (size_t)n_orig * (size_t)(max_lag + 1);
    const bool tiled = !(tiled_env && atoi(tiled_env) == 0) && n_snap >= 2 && max_lag >= 1 && dmat_bytes <= ((size_t)2 << 30);
    DevBuf dmat;
    if (tiled) LJMD_CUDA_CHECK(cudaMalloc(&dmat.p, dmat_bytes));
    EventPair ev;
    LJMD_CUDA_CHECK(cudaEventCreate(&ev.e0));
    LJMD_CUDA_CHECK(cudaEventCreate(&ev.e1));
    LJMD_CUDA_CHECK(cudaEventRecord(ev.e0, 0));
    if (tiled) {
        const dim3 tgrid((unsigned)((max_lag + 1 + CT_WL * CT_TL - 1) / (CT_WL * CT_TL)),
                         (unsigned)((n_orig + CT_WO * CT_TO - 1) / (CT_WO * CT_TO)));
        if (tgrid.y > 65535u) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_time_correlation(): too many origins for the tiled kernel.");
        double *dm = dmat.as<double>();
        if (kind == LJMD_CORR_MSD) {
            if (origin_stride == 1) time_corr_tiled_kernel<0, true><<<tgrid, CORR_BS>>>(a, dm, n_orig);
            else time_corr_tiled_kernel<0, false><<<tgrid, CORR_BS>>>(a, dm, n_orig);
        } else {
            if (origin_stride == 1) time_corr_tiled_kernel<1, true><<<tgrid, CORR_BS>>>(a, dm, n_orig);
            else time_corr_tiled_kernel<1, false><<<tgrid, CORR_BS>>>(a, dm, n_orig);
        }
        time_corr_reduce_kernel<<<(unsigned)((max_lag + CORR_BS) / CORR_BS), CORR_BS>>>(a, dm, n_orig);
    } else if (kind == LJMD_CORR_MSD) {
        time_corr_kernel<0><<<grid, CORR_BS>>>(a);
    } else {
        time_corr_kernel<1><<<grid, CORR_BS>>>(a);
    }
    LJMD_CUDA_CHECK(cudaGetLastError());
    LJMD_CUDA_CHECK(cudaEventRecord(ev.e1, 0));
    LJMD_CUDA_CHECK(cudaDeviceSynchronize());
    if (kernel_ms) {
        float ms = 0.f;
        LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        *kernel_ms = ms;
    }
    LJMD_CUDA_CHECK(cudaMemcpy(out, dout.p, sizeof(double) * (max_lag + 1), cudaMemcpyDeviceToHost));
    return LJMD_OK;
}
