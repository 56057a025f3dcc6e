// ljmd_api.cu -- the C ABI of libljmd.so (include/ljmd.h): parameter handling, contexts,
// host<->device staging, CUDA-graph step driver and the stateless drop-in pair that replaces
// the reference's compute_lj_potential_energy / verlet_step
// (/root/reference/scripts/physics/lj_potential_energy.f90:46, physics/verlet.f90:41).
//
// There is no CPU fallback in this library: every compute entry point needs a CUDA device.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <new>
#include <numeric>
#include <vector>

#include "ljmd_internal.h"

namespace ljmd {

static thread_local char g_err[512] = {0};

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }

// The same guards, in the same order, as lj_potential_energy.f90:77-81 (+ md_types.f90:152).
static int check_params(const ljmd_params *p, const char *who)
{
    if (!p) return set_error(LJMD_ERR_NULL, "%s(): params is NULL.", who);
    if (p->n <= 0) return set_error(LJMD_ERR_N, "%s(): params%%n must be > 0.", who);
    if (!(p->box_length > 0.0))
        return set_error(LJMD_ERR_BOX, "%s(): params%%box_length must be > 0.", who);
    if (!(p->volume > 0.0))
        return set_error(LJMD_ERR_VOLUME, "%s(): params%%volume must be > 0.", who);
    if (!(p->rc > 0.0)) return set_error(LJMD_ERR_RC, "%s(): params%%rc must be > 0.", who);
    if (!(p->rc_square > 0.0))
        return set_error(LJMD_ERR_RC_SQUARE, "%s(): params%%rc_square must be > 0.", who);
    if (p->rc >= 0.5 * p->box_length)
        return set_error(LJMD_ERR_RC_HALF_BOX,
                         "%s(): rc (cutoff_radius) must be < L/2 (minimum image convention).", who);
    return LJMD_OK;
}

// Tail constants, lj_potential_energy.f90:205-212, in the reference's operation order
// (integer powers are repeated multiplications; pi from md_types.f90:22).
static void tail_constants(const ljmd_params &p, double *tu, double *tdu, double *tddu)
{
    const double pi = 3.1415926535897932384626433832795;
    volatile double npd = (double)p.n;
    volatile double rc3 = (p.rc * p.rc) * p.rc;
    volatile double rc6 = rc3 * rc3;
    volatile double tail_factor = 8.0 * pi * (npd * npd) / (p.volume * rc3);
    *tu = tail_factor * ((1.0 / (3.0 * rc6)) - 1.0) / 3.0;
    *tdu = 2.0 * tail_factor * (-2.0 / (3.0 * rc6) + 1.0);
    *tddu = 2.0 * tail_factor * (26.0 / (3.0 * rc6) - 7.0);
}

static void refresh_constants(ljmd_ctx *c)
{
    SimConst &s = c->sc;
    s.n = c->p.n;
    s.R = c->R;
    s.L = c->p.box_length;
    s.invL = c->p.inv_box_length;
    s.rc2 = c->p.rc_square;
    s.dt = c->p.dt;
    s.dt_half = c->p.dt_half;
    s.dt_square_half = c->p.dt_square_half;
    tail_constants(c->p, &s.tail_u, &s.tail_du, &s.tail_ddu);
}

static void drop_graphs(ljmd_ctx *c)
{
    if (c->graph_chunk) { cudaGraphExecDestroy(c->graph_chunk); c->graph_chunk = nullptr; }
    if (c->graph_compat) { cudaGraphExecDestroy(c->graph_compat); c->graph_compat = nullptr; }
    for (int k = 0; k < 16; ++k)
        if (c->graph_tail[k]) { cudaGraphExecDestroy(c->graph_tail[k]); c->graph_tail[k] = nullptr; }
    for (int k = 0; k < 33; ++k)
        if (c->graph_slab[k]) { cudaGraphExecDestroy(c->graph_slab[k]); c->graph_slab[k] = nullptr; }
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int dev)
    {
        LJMD_CUDA_CHECK(cudaGetDevice(&prev));
        if (dev != prev) { LJMD_CUDA_CHECK(cudaSetDevice(dev)); active = true; }
        return LJMD_OK;
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------
// Spatial order of the sorted all-pairs path (lj_allpairs.cu: lj_pairs_sorted_kernel): recursive
// bisection at the median of the widest extent, left part a multiple of 128 atoms (32 below 128),
// down to leaves of 32 atoms = the j-chunks; inside a leaf atoms keep their original order.
// Deterministic (ties broken by atom index).  Any order is CORRECT -- the kernel tests every pair
// and verifies its per-tile assumptions -- this one only makes the tiles compact.
// ------------------------------------------------------------------------------------------
static void kd_order_rec(int *idx, int m, const double *x, const double *y, const double *z)
{
    if (m <= 32) { std::sort(idx, idx + m); return; }
    const int unit = m > 128 ? 128 : 32;
    const int nu = (m + unit - 1) / unit;
    int nl = ((nu + 1) / 2) * unit;
    if (nl > m - 1) nl = m - 1;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    const double *c3[3] = {x, y, z};
    for (int k = 0; k < m; ++k)
        for (int d = 0; d < 3; ++d) {
            const double v = c3[d][idx[k]];
            lo[d] = std::min(lo[d], v);
            hi[d] = std::max(hi[d], v);
        }
    int d = 0;
    if (hi[1] - lo[1] > hi[d] - lo[d]) d = 1;
    if (hi[2] - lo[2] > hi[d] - lo[d]) d = 2;
    const double *cd = c3[d];
    std::nth_element(idx, idx + nl, idx + m, [cd](int a, int b) { return cd[a] < cd[b] || (cd[a] == cd[b] && a < b); });
    kd_order_rec(idx, nl, x, y, z);
    kd_order_rec(idx + nl, m - nl, x, y, z);
}

static int compute_sort_perm(ljmd_ctx *c, const double *x, const double *y, const double *z)
{
    const int n = c->sc.n;
    if (!c->pw.perm) {
        c->pw.perm = (int *)malloc(sizeof(int) * (size_t)n);
        if (!c->pw.perm) return set_error(LJMD_ERR_NOMEM, "out of host memory (sort permutation).");
    }
    std::iota(c->pw.perm, c->pw.perm + n, 0);
    bool finite = true;
    for (int k = 0; k < n && finite; ++k) finite = std::isfinite(x[k]) && std::isfinite(y[k]) && std::isfinite(z[k]);
    if (finite) kd_order_rec(c->pw.perm, n, x, y, z); // NaN / Inf input: keep the caller's order (the reference does not guard either)
    c->pw.steps_since_sort = 0;
    return LJMD_OK;
}

// host array (caller's order) -> pinned stage (device order) and back
static void stage_in(const ljmd_ctx *c, double *dst, const double *src)
{
    const size_t n = c->nR;
    if (const int *pm = c->pw.perm) for (size_t k = 0; k < n; ++k) dst[k] = src[pm[k]];
    else memcpy(dst, src, sizeof(double) * n);
}
static void stage_out(const ljmd_ctx *c, double *dst, const double *src)
{
    const size_t n = c->nR;
    if (const int *pm = c->pw.perm) for (size_t k = 0; k < n; ++k) dst[pm[k]] = src[k];
    else memcpy(dst, src, sizeof(double) * n);
}

// Re-sort of a spatially sorted all-pairs context between two steps (positions in rbuf[0]): diffusion slowly
// loosens the chunk boxes the kernel's per-tile tests rely on (it stays correct, more tiles fall back to the
// general path).  Host-side like the first sort: D2H of the state, new order from the current positions, H2D.
// The resident order is not observable (upload / download translate), the sums change order: results move by
// rounding only, and WHEN it happens depends on the step count since the last complete upload alone.
static int resort_allpairs(ljmd_ctx *c)
{
    const size_t n = c->nR, nb = sizeof(double) * n;
    const int narr = c->st.rux ? 12 : 9;
    std::vector<double> host((size_t)narr * n), tmp(n);
    LJMD_CUDA_CHECK(cudaMemcpyAsync(host.data(), c->state_block, 9 * nb, cudaMemcpyDeviceToHost, c->stream));
    if (c->st.rux) LJMD_CUDA_CHECK(cudaMemcpyAsync(host.data() + 9 * n, c->st.rux, 3 * nb, cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<int> idx(n), newperm(n);
    std::iota(idx.begin(), idx.end(), 0);
    kd_order_rec(idx.data(), (int)n, host.data(), host.data() + n, host.data() + 2 * n); // new slot k <- old slot idx[k]
    for (size_t k = 0; k < n; ++k) newperm[k] = c->pw.perm ? c->pw.perm[idx[k]] : idx[k];
    if (!c->pw.perm) {
        c->pw.perm = (int *)malloc(sizeof(int) * n);
        if (!c->pw.perm) return set_error(LJMD_ERR_NOMEM, "out of host memory (sort permutation).");
    }
    memcpy(c->pw.perm, newperm.data(), sizeof(int) * n);
    for (int a = 0; a < narr; ++a) {
        double *arr = host.data() + (size_t)a * n;
        for (size_t k = 0; k < n; ++k) tmp[k] = arr[idx[k]];
        memcpy(arr, tmp.data(), nb);
    }
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->state_block, host.data(), 9 * nb, cudaMemcpyHostToDevice, c->stream));
    if (c->st.rux) LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.rux, host.data() + 9 * n, 3 * nb, cudaMemcpyHostToDevice, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // `host` goes out of scope
    c->pw.steps_since_sort = 0;
    return LJMD_OK;
}

static long long sort_interval()
{
    // steps between re-sorts; <= 0: never.  A re-sort costs ~3 ms of host work at N = 32k (6 us per step at the
    // default) and a freshly sorted liquid runs the force kernel 7 % faster than the order of a lattice start
    // (0.432 vs 0.466 ms, profiles/r2_bench_cfg3_allpairs_resort50.json).
    const char *e = getenv("LJMD_SORT_INTERVAL");
    return e ? atoll(e) : 500;
}

static int enqueue_force(ljmd_ctx *c, int buf)
{
    if (c->method == LJMD_METHOD_CELLS) return launch_cells_force(c, buf);
    return launch_pairs(c, buf);
}

static int kernels_per_step(const ljmd_ctx *c)
{
    if (c->method == LJMD_METHOD_CELLS) return c->cw.kernels_per_force + 1;
    return c->pw.sorted ? 3 : 2; // (chunk boxes +) force + finalize
}

// One velocity-Verlet step.  On entry rbuf[prev] holds r(t) and rbuf[1-prev] already holds the
// drifted + wrapped r(t+dt).  fin_mode 1 leaves r(t+2dt) in rbuf[prev]; fin_mode 2 copies r(t+dt)
// into rbuf[prev] instead (used for the last step of an odd-length run so that rbuf[0] is current).
static int enqueue_step(ljmd_ctx *c, int prev, int fin_mode)
{
    const int now = 1 - prev;
    LJMD_TRY(enqueue_force(c, now));
    LJMD_TRY(launch_finalize(c, fin_mode, prev, now, prev));
    return LJMD_OK;
}

static int capture_graph(ljmd_ctx *c, int nsteps, cudaGraphExec_t *out)
{
    cudaGraph_t g = nullptr;
    LJMD_CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    int rc = LJMD_OK;
    for (int s = 0; s < nsteps && rc == LJMD_OK; ++s) rc = enqueue_step(c, s & 1, 1);
    c->capturing = false;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc != LJMD_OK) { if (g) cudaGraphDestroy(g); return rc; }
    LJMD_CUDA_CHECK(e);
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    LJMD_CUDA_CHECK(e);
    return LJMD_OK;
}

static int ensure_stage(ljmd_ctx *c)
{
    if (!c->h_stage) {
        const size_t bytes = sizeof(double) * (9 * c->nR + 4 * (size_t)c->R);
        LJMD_CUDA_CHECK(cudaMallocHost(&c->h_stage, bytes));
    }
    return LJMD_OK;
}

static int set_ctl(ljmd_ctx *c, long long cap, double *series)
{
    StepCtl h;
    h.step = 0;
    h.series_cap = cap;
    h.series = series;
    // small synchronous-with-respect-to-stream copy from pageable memory (24 bytes)
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->fw.ctl, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

// nsteps velocity-Verlet steps.  Invariant outside this function: rbuf[0] (the state block's
// rx..rz) holds the current positions.
static int run_steps_nlist(ljmd_ctx *c, int nsteps)
{
    LJMD_TRY(nlist_enqueue_drift(c)); // r(t+dt) of the first step -> rb[1-cur]
    int done = 0;
    if (c->timing) {
        for (; done < nsteps; ++done) {
            LJMD_TRY(nlist_enqueue_rebuild(c, 1));
            LJMD_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
            LJMD_TRY(nlist_enqueue_step_kernel(c));
            LJMD_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
            LJMD_CUDA_CHECK(cudaEventSynchronize(c->ev1));
            float ms = 0.f;
            LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            c->force_ms += ms;
            c->force_evals += 1;
        }
        return LJMD_OK;
    }
    const int G = 32; // even: a chunk leaves NlCtl::cur unchanged
    if (!getenv("LJMD_NL_NOGRAPH")) {
        for (; done + G <= nsteps; done += G) {
            LJMD_CUDA_CHECK(cudaGraphLaunch(c->graph_chunk, c->stream));
            c->launches += (int64_t)G * 2; // condition kernel + step kernel; rebuilds are added at the next sync
        }
        const int tail = (nsteps - done) & ~1;
        if (tail >= 2) {
            LJMD_CUDA_CHECK(cudaGraphLaunch(c->graph_tail[tail / 2], c->stream));
            c->launches += (int64_t)tail * 2;
            done += tail;
        }
    }
    for (; done < nsteps; ++done) LJMD_TRY(nlist_enqueue_step(c));
    return LJMD_OK;
}

// Captures the graphs a run of nsteps needs (32-step chunks plus ONE even-length tail graph, cached by
// length) BEFORE anything is enqueued, so that capture / instantiation never sits between two launches.
static int prepare_graphs(ljmd_ctx *c, int nsteps)
{
    if (c->timing || nsteps < 2) return LJMD_OK;
    const bool nl = c->method == LJMD_METHOD_NLIST;
    if (nl && getenv("LJMD_NL_NOGRAPH")) return LJMD_OK;
    const int G = 32;
    if (nsteps >= G && !c->graph_chunk) {
        LJMD_TRY(nl ? nlist_capture_graph(c, G, &c->graph_chunk) : capture_graph(c, G, &c->graph_chunk));
        c->graph_chunk_steps = G;
    }
    const int tail = (nsteps % G) & ~1;
    if (tail >= 2 && !c->graph_tail[tail / 2])
        LJMD_TRY(nl ? nlist_capture_graph(c, tail, &c->graph_tail[tail / 2]) : capture_graph(c, tail, &c->graph_tail[tail / 2]));
    return LJMD_OK;
}

static int run_steps_segment(ljmd_ctx *c, int nsteps);
static int prepare_graphs(ljmd_ctx *c, int nsteps);

static int run_steps(ljmd_ctx *c, int nsteps)
{
    if (nsteps <= 0) return LJMD_OK;
    if (c->method == LJMD_METHOD_NLIST) return run_steps_nlist(c, nsteps);
    const long long iv = sort_interval();
    if (c->method != LJMD_METHOD_ALLPAIRS || !c->pw.sorted || iv <= 0 || c->R != 1) return run_steps_segment(c, nsteps);
    // sorted all-pairs: segments that end at the re-sort points (multiples of the interval since the last complete upload)
    int left = nsteps;
    while (left > 0) {
        const long long room = iv - c->pw.steps_since_sort;
        const int todo = (int)std::min<long long>(left, room > 0 ? room : iv);
        if (todo != nsteps) LJMD_TRY(prepare_graphs(c, todo)); // the caller prepared the graphs of an unsplit run
        LJMD_TRY(run_steps_segment(c, todo));
        c->pw.steps_since_sort += todo;
        left -= todo;
        if (c->pw.steps_since_sort >= iv) LJMD_TRY(resort_allpairs(c));
    }
    return LJMD_OK;
}

static int run_steps_segment(ljmd_ctx *c, int nsteps)
{
    LJMD_TRY(launch_drift(c, 0, 1)); // r(t+dt) of the first step
    int done = 0;
    if (c->timing) {
        for (; done < nsteps; ++done) {
            const int prev = done & 1, last = (done == nsteps - 1);
            LJMD_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
            LJMD_TRY(enqueue_force(c, 1 - prev));
            LJMD_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
            LJMD_TRY(launch_finalize(c, (last && prev == 0) ? 2 : 1, prev, 1 - prev, prev));
            LJMD_CUDA_CHECK(cudaEventSynchronize(c->ev1));
            float ms = 0.f;
            LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            c->force_ms += ms;
            c->force_evals += 1;
        }
        return LJMD_OK;
    }
    const int G = 32; // even, so a chunk starts and ends with rbuf[0] = r(t)
    for (; done + G <= nsteps; done += G) {
        LJMD_CUDA_CHECK(cudaGraphLaunch(c->graph_chunk, c->stream));
        c->launches += (int64_t)G * kernels_per_step(c);
    }
    const int tail = (nsteps - done) & ~1; // even, for the same reason
    if (tail >= 2) {
        LJMD_CUDA_CHECK(cudaGraphLaunch(c->graph_tail[tail / 2], c->stream));
        c->launches += (int64_t)tail * kernels_per_step(c);
        done += tail;
    }
    for (; done < nsteps; ++done) {
        const int prev = done & 1, last = (done == nsteps - 1);
        LJMD_TRY(enqueue_step(c, prev, (last && prev == 0) ? 2 : 1));
    }
    return LJMD_OK;
}

static int fetch_obs(ljmd_ctx *c, double *dst, int per_rep, const int *idx)
{
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_obs, c->fw.obs, sizeof(double) * 4 * c->R,
                                    cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (dst)
        for (int r = 0; r < c->R; ++r)
            for (int k = 0; k < per_rep; ++k) dst[(size_t)r * per_rep + k] = c->h_obs[4 * r + idx[k]];
    if (c->method == LJMD_METHOD_NLIST) {
        LJMD_TRY(nlist_check_overflow(c));
        // list builds that ran inside graph IF nodes: 9 kernels each
        if (c->graph_chunk) c->launches += 9 * (c->nw.rebuilds_host - c->nw.rebuilds_counted);
        c->nw.rebuilds_counted = c->nw.rebuilds_host;
    }
    return LJMD_OK;
}

} // namespace ljmd

using namespace ljmd;

// ==========================================================================================
// Library / diagnostics
// ==========================================================================================
extern "C" int ljmd_version(void) { return LJMD_VERSION; }
extern "C" const char *ljmd_last_error(void) { return g_err; }
extern "C" int ljmd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ==========================================================================================
// init_params + compute_derived_params, base/md_types.f90:105-169
// ==========================================================================================
extern "C" int ljmd_params_init(ljmd_params *p, int32_t n, double box_length, double dt,
                                double rc, int32_t num_cells)
{
    clear_error();
    if (!p) return set_error(LJMD_ERR_NULL, "ljmd_params_init(): params is NULL.");
    memset(p, 0, sizeof(*p));
    p->n = n;
    p->box_length = box_length;
    p->dt = dt;
    p->rc = rc;
    p->num_cells = num_cells;
    if (p->box_length > 0.0) {
        p->inv_box_length = 1.0 / p->box_length;                     // :137
        volatile double l2 = p->box_length * p->box_length;
        p->volume = l2 * p->box_length;                              // :138  L**3
        if (p->n > 0) p->density = (double)p->n / p->volume;         // :141
    } else {
        return set_error(LJMD_ERR_BOX, "compute_derived_params(): box_length must be > 0.");
    }
    if (p->rc > 0.0) {
        p->rc_square = p->rc * p->rc;                                // :148
    } else {
        return set_error(LJMD_ERR_RC, "compute_derived_params(): rc (cutoff_radius) must be > 0.");
    }
    if (p->rc >= 0.5 * p->box_length)                                // :152
        return set_error(LJMD_ERR_RC_HALF_BOX,
                         "compute_derived_params(): rc (cutoff_radius) must be < L/2 (minimum image convention).");
    if (p->dt > 0.0) {
        p->dt_half = 0.5 * p->dt;                                    // :158
        p->dt_square_half = p->dt_half * p->dt;                      // :159
    } else {
        return set_error(LJMD_ERR_DT, "compute_derived_params(): dt must be > 0.");
    }
    if (p->length_scale > 0.0) p->inv_length_scale = 1.0 / p->length_scale;
    return LJMD_OK;
}

// ==========================================================================================
// Contexts
// ==========================================================================================
extern "C" int ljmd_ctx_destroy(ljmd_ctx *c)
{
    if (!c) return LJMD_OK;
    DeviceGuard g;
    g.enter(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    cellwork_free(c);
    nlist_free(c);
    cudaFree(c->state_block);
    cudaFree(c->pw.fpart);
    cudaFree(c->pw.spart);
    cudaFree(c->pw.fJ);
    cudaFree(c->pw.items);
    cudaFree(c->pw.row_splits);
    cudaFree(c->pw.dbg);
    cudaFree(c->pw.cbox);
    cudaFree(c->pw.near);
    cudaFree(c->pw.work);
    free(c->pw.perm);
    cudaFree(c->fw.kpart);
    cudaFree(c->fw.counter);
    cudaFree(c->fw.ctl);
    cudaFree(c->fw.pair_counts);
    cudaFree(c->d_series);
    if (c->h_obs) cudaFreeHost(c->h_obs);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_s0) cudaEventDestroy(c->ev_s0);
    if (c->ev_s1) cudaEventDestroy(c->ev_s1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    cudaGetLastError();
    delete c;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_create(const ljmd_params *params, const ljmd_options *opts,
                               ljmd_ctx **out)
{
    clear_error();
    if (!out) return set_error(LJMD_ERR_NULL, "ljmd_ctx_create(): out is NULL.");
    *out = nullptr;
    LJMD_TRY(check_params(params, "ljmd_ctx_create"));
    ljmd_options o = {LJMD_METHOD_AUTO, -1, 1, 0};
    if (opts) o = *opts;
    if (o.nreplicas < 1) return set_error(LJMD_ERR_ARG, "ljmd_ctx_create(): nreplicas must be >= 1.");
    if (o.method < LJMD_METHOD_AUTO || o.method > LJMD_METHOD_NLIST)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_create(): unknown method %d.", o.method);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_error(LJMD_ERR_CUDA, "ljmd_ctx_create(): no CUDA device available (this library has no CPU fallback).");
    }
    int dev = o.device;
    if (dev < 0) LJMD_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= ndev) return set_error(LJMD_ERR_ARG, "ljmd_ctx_create(): device %d out of range.", dev);

    ljmd_ctx *c = new (std::nothrow) ljmd_ctx();
    if (!c) return set_error(LJMD_ERR_NOMEM, "ljmd_ctx_create(): out of host memory.");
    memset(c, 0, sizeof(*c));
    c->p = *params;
    c->o = o;
    c->device = dev;
    c->R = o.nreplicas;
    c->nR = (size_t)params->n * c->R;

    const int nc = cells_per_edge_host(*params);
    const int nc_nl = nlist_cells_per_edge_host(*params, nlist_default_skin());
    int method = o.method;
    if (method == LJMD_METHOD_AUTO)
        method = (nc_nl >= 4 && params->n >= 8192 && o.nreplicas == 1) ? LJMD_METHOD_NLIST : LJMD_METHOD_ALLPAIRS;
    if (method == LJMD_METHOD_CELLS && nc < 3) {
        delete c;
        return set_error(LJMD_ERR_UNSUPPORTED,
                         "ljmd_ctx_create(): cell list needs floor(L/rc) >= 3 (got %d).", nc);
    }
    if (method == LJMD_METHOD_NLIST && nc_nl < 3) {
        delete c;
        return set_error(LJMD_ERR_UNSUPPORTED,
                         "ljmd_ctx_create(): neighbour list needs floor(L/(rc+skin)) >= 3 (got %d).", nc_nl);
    }
    c->method = method;

#define CREATE_CHECK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            int rc__ = set_error(LJMD_ERR_CUDA, "%s failed: %s", #expr,                     \
                                 cudaGetErrorString(err__));                                \
            ljmd_ctx_destroy(c);                                                            \
            return rc__;                                                                    \
        }                                                                                   \
    } while (0)
#define CREATE_TRY(expr)                                                                    \
    do {                                                                                    \
        int rc__ = (expr);                                                                  \
        if (rc__ != LJMD_OK) { ljmd_ctx_destroy(c); return rc__; }                          \
    } while (0)

    DeviceGuard g;
    CREATE_TRY(g.enter(dev));
    CREATE_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = c->stream;
    CREATE_CHECK(cudaEventCreate(&c->ev0));
    CREATE_CHECK(cudaEventCreate(&c->ev1));
    CREATE_CHECK(cudaEventCreate(&c->ev_s0));
    CREATE_CHECK(cudaEventCreate(&c->ev_s1));
    refresh_constants(c);

    // state block: 9 arrays | obs[R][4] | alt positions (3 arrays) | (ru: 3 arrays)
    const bool track = (o.flags & LJMD_FLAG_TRACK_UNWRAPPED) != 0;
    const size_t ndoubles = 12 * c->nR + 4 * (size_t)c->R + (track ? 3 * c->nR : 0);
    CREATE_CHECK(cudaMalloc(&c->state_block, sizeof(double) * ndoubles));
    CREATE_CHECK(cudaMemsetAsync(c->state_block, 0, sizeof(double) * ndoubles, c->stream)); // zero_state, md_types.f90:222-230
    double *b = c->state_block;
    c->st.rx = b + 0 * c->nR; c->st.ry = b + 1 * c->nR; c->st.rz = b + 2 * c->nR;
    c->st.vx = b + 3 * c->nR; c->st.vy = b + 4 * c->nR; c->st.vz = b + 5 * c->nR;
    c->st.ax = b + 6 * c->nR; c->st.ay = b + 7 * c->nR; c->st.az = b + 8 * c->nR;
    c->fw.obs = b + 9 * c->nR;
    c->rbuf[0] = b;
    c->rbuf[1] = b + 9 * c->nR + 4 * (size_t)c->R;
    if (track) {
        double *u = b + 12 * c->nR + 4 * (size_t)c->R;
        c->st.rux = u; c->st.ruy = u + c->nR; c->st.ruz = u + 2 * c->nR;
    }

    c->fw.block = 256;
    // small systems: 8 thread groups share each atom's partial sum (many partials, few atoms)
    c->fw.groups = (method == LJMD_METHOD_ALLPAIRS && params->n <= 65536) ? 8 : 1;
    c->fw.nblk = (params->n + (256 / c->fw.groups) - 1) / (256 / c->fw.groups);
    CREATE_CHECK(cudaMalloc(&c->fw.kpart, sizeof(double) * 6 * (size_t)c->fw.nblk * c->R)); // per CTA: 3 E_kin + 3 pair-sum partials
    CREATE_CHECK(cudaMalloc(&c->fw.counter, sizeof(unsigned) * (c->R + 1)));
    CREATE_CHECK(cudaMemsetAsync(c->fw.counter, 0, sizeof(unsigned) * (c->R + 1), c->stream));
    CREATE_CHECK(cudaMalloc(&c->fw.ctl, sizeof(StepCtl)));
    CREATE_CHECK(cudaMemsetAsync(c->fw.ctl, 0, sizeof(StepCtl), c->stream));
    CREATE_CHECK(cudaMallocHost(&c->h_obs, sizeof(double) * 4 * (size_t)c->R + 64));

    if (method == LJMD_METHOD_CELLS) CREATE_TRY(cellwork_plan(c));
    else if (method == LJMD_METHOD_NLIST) CREATE_TRY(nlist_plan(c));
    else CREATE_TRY(pairwork_plan(c));
    CREATE_CHECK(cudaStreamSynchronize(c->stream));
#undef CREATE_CHECK
#undef CREATE_TRY
    *out = c;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_upload(ljmd_ctx *c, const double *rx, const double *ry, const double *rz,
                               const double *vx, const double *vy, const double *vz,
                               const double *ax, const double *ay, const double *az)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_upload(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(ensure_stage(c));
    // the neighbour-list path keeps the state cell-sorted: back to original order first
    if (c->method == LJMD_METHOD_NLIST) LJMD_TRY(nlist_canonicalize(c));
    const double *src[9] = {rx, ry, rz, vx, vy, vz, ax, ay, az};
    const size_t bytes = sizeof(double) * c->nR;
    // sorted all-pairs path: a COMPLETE upload defines a new spatial order; a partial one keeps the
    // current order (the arrays it leaves alone are already on the device in that order)
    if (c->method == LJMD_METHOD_ALLPAIRS && c->pw.sorted && rx && ry && rz && vx && vy && vz && ax && ay && az)
        LJMD_TRY(compute_sort_perm(c, rx, ry, rz));
    // pack contiguous runs of non-NULL arrays into the pinned stage, one H2D per run
    int k = 0;
    while (k < 9) {
        if (!src[k]) { ++k; continue; }
        int e = k;
        while (e < 9 && src[e]) { stage_in(c, c->h_stage + (size_t)e * c->nR, src[e]); ++e; }
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->state_block + (size_t)k * c->nR,
                                        c->h_stage + (size_t)k * c->nR, bytes * (e - k),
                                        cudaMemcpyHostToDevice, c->stream));
        k = e;
    }
    if (c->st.rux && (rx || ry || rz)) { // ru(t=0) = r(t=0), md_simulation_program.f90:229-231
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.rux, c->st.rx, 3 * bytes, cudaMemcpyDeviceToDevice,
                                        c->stream));
    }
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_download(ljmd_ctx *c, double *rx, double *ry, double *rz, double *vx,
                                 double *vy, double *vz, double *ax, double *ay, double *az)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_download(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(ensure_stage(c));
    double *dst[9] = {rx, ry, rz, vx, vy, vz, ax, ay, az};
    const size_t bytes = sizeof(double) * c->nR;
    const double *dev_src = c->state_block;
    if (c->method == LJMD_METHOD_NLIST) { // un-sort into the scratch block first
        LJMD_TRY(nlist_gather_orig(c));
        dev_src = c->nw.alt;
    }
    int k = 0;
    while (k < 9) {
        if (!dst[k]) { ++k; continue; }
        int e = k;
        while (e < 9 && dst[e]) ++e;
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + (size_t)k * c->nR,
                                        dev_src + (size_t)k * c->nR, bytes * (e - k),
                                        cudaMemcpyDeviceToHost, c->stream));
        k = e;
    }
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (k = 0; k < 9; ++k)
        if (dst[k]) stage_out(c, dst[k], c->h_stage + (size_t)k * c->nR);
    return LJMD_OK;
}

extern "C" int ljmd_ctx_download_unwrapped(ljmd_ctx *c, double *rux, double *ruy, double *ruz)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_download_unwrapped(): ctx is NULL.");
    if (!c->st.rux)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_download_unwrapped(): context was created without LJMD_FLAG_TRACK_UNWRAPPED.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    const size_t bytes = sizeof(double) * c->nR;
    double *dst[3] = {rux, ruy, ruz};
    double *srcs[3] = {c->st.rux, c->st.ruy, c->st.ruz};
    if (c->method == LJMD_METHOD_NLIST) {
        LJMD_TRY(nlist_gather_orig(c));
        for (int k = 0; k < 3; ++k) srcs[k] = c->nw.alt + (size_t)(9 + k) * c->nR;
    }
    if (c->method == LJMD_METHOD_ALLPAIRS && c->pw.perm) { // device order -> caller's order through the stage
        LJMD_TRY(ensure_stage(c));
        for (int k = 0; k < 3; ++k)
            if (dst[k]) LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + (size_t)k * c->nR, srcs[k], bytes, cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 3; ++k)
            if (dst[k]) stage_out(c, dst[k], c->h_stage + (size_t)k * c->nR);
        return LJMD_OK;
    }
    for (int k = 0; k < 3; ++k)
        if (dst[k]) LJMD_CUDA_CHECK(cudaMemcpyAsync(dst[k], srcs[k], bytes, cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_compute_forces(ljmd_ctx *c, double *obs3)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_compute_forces(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    if (c->method == LJMD_METHOD_NLIST) {
        LJMD_TRY(nlist_enqueue_forces(c));
    } else {
        LJMD_TRY(enqueue_force(c, 0));
        LJMD_TRY(launch_finalize(c, 0, 0, 0, 0));
    }
    const int idx[3] = {0, 2, 3};
    return fetch_obs(c, obs3, 3, idx);
}

extern "C" int ljmd_ctx_step(ljmd_ctx *c, int32_t nsteps, double *obs4)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_step(): ctx is NULL.");
    if (nsteps < 0) return set_error(LJMD_ERR_ARG, "ljmd_ctx_step(): nsteps must be >= 0.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(prepare_graphs(c, nsteps));
    LJMD_CUDA_CHECK(cudaEventRecord(c->ev_s0, c->stream));
    LJMD_TRY(run_steps(c, nsteps));
    LJMD_CUDA_CHECK(cudaEventRecord(c->ev_s1, c->stream));
    const int idx[4] = {0, 1, 2, 3};
    LJMD_TRY(fetch_obs(c, obs4, 4, idx));
    float ms = 0.f;
    LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev_s0, c->ev_s1));
    c->last_step_ms = ms;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_step_series(ljmd_ctx *c, int32_t nsteps, double *series)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_step_series(): ctx is NULL.");
    if (nsteps < 0) return set_error(LJMD_ERR_ARG, "ljmd_ctx_step_series(): nsteps must be >= 0.");
    if (!series) return set_error(LJMD_ERR_NULL, "ljmd_ctx_step_series(): series is NULL.");
    if (nsteps == 0) return LJMD_OK;
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    if (c->series_cap < nsteps) {
        cudaFree(c->d_series);
        c->d_series = nullptr;
        c->series_cap = 0;
        LJMD_CUDA_CHECK(cudaMalloc(&c->d_series, sizeof(double) * 4 * (size_t)c->R * nsteps));
        c->series_cap = nsteps;
    }
    LJMD_TRY(set_ctl(c, nsteps, c->d_series));
    int rc = prepare_graphs(c, nsteps);
    if (rc == LJMD_OK) rc = run_steps(c, nsteps);
    if (rc == LJMD_OK) {
        cudaError_t e = cudaMemcpyAsync(series, c->d_series, sizeof(double) * 4 * (size_t)c->R * nsteps,
                                        cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "series download failed: %s", cudaGetErrorString(e));
    }
    int rc2 = set_ctl(c, 0, nullptr);
    return rc != LJMD_OK ? rc : rc2;
}

// Arms (cap > 0) / disarms (cap == 0) the per-step observable recorder: every following step stores
// its 4 observables per replica at series[step][replica][4], counting steps from this call.
extern "C" int ljmd_ctx_series_arm(ljmd_ctx *c, int32_t cap)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_series_arm(): ctx is NULL.");
    if (cap < 0) return set_error(LJMD_ERR_ARG, "ljmd_ctx_series_arm(): cap must be >= 0.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    if (cap == 0) return set_ctl(c, 0, nullptr);
    if (c->series_cap < cap) {
        cudaFree(c->d_series);
        c->d_series = nullptr;
        c->series_cap = 0;
        LJMD_CUDA_CHECK(cudaMalloc(&c->d_series, sizeof(double) * 4 * (size_t)c->R * cap));
        c->series_cap = cap;
    }
    return set_ctl(c, cap, c->d_series);
}

// Copies the first nsteps recorded steps to the host.
extern "C" int ljmd_ctx_series_fetch(ljmd_ctx *c, int32_t nsteps, double *series)
{
    clear_error();
    if (!c || !series) return set_error(LJMD_ERR_NULL, "ljmd_ctx_series_fetch(): NULL argument.");
    if (nsteps < 0 || nsteps > c->series_cap) return set_error(LJMD_ERR_ARG, "ljmd_ctx_series_fetch(): nsteps out of range.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_CUDA_CHECK(cudaMemcpyAsync(series, c->d_series, sizeof(double) * 4 * (size_t)c->R * nsteps, cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_kinetic_energy(ljmd_ctx *c, double *ekin)
{
    clear_error();
    if (!c || !ekin) return set_error(LJMD_ERR_NULL, "ljmd_ctx_kinetic_energy(): NULL argument.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    // reuse kpart as scratch output (R doubles fit: nblk*3*R >= R)
    LJMD_TRY(launch_ekin_driver(c, c->fw.kpart));
    LJMD_CUDA_CHECK(cudaMemcpyAsync(ekin, c->fw.kpart, sizeof(double) * c->R, cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_last_step_ms(const ljmd_ctx *c, double *ms)
{
    if (!c || !ms) return set_error(LJMD_ERR_NULL, "ljmd_ctx_last_step_ms(): NULL argument.");
    *ms = c->last_step_ms;
    return LJMD_OK;
}

// Tuning aid: copies the per-warp {start, end, smid, items} records of the last independent-warp
// force launch (needs LJMD_SYM_DEBUG=1 at context creation).  Returns the number of warps.
extern "C" int ljmd_ctx_debug_warp_times(ljmd_ctx *c, long long *out, int max_warps)
{
    if (!c || !c->pw.dbg) return 0;
    const int nw = c->pw.n_ctas * c->pw.cta_warps;
    const int m = nw < max_warps ? nw : max_warps;
    cudaStreamSynchronize(c->stream);
    if (out && cudaMemcpy(out, c->pw.dbg, sizeof(long long) * 4 * (size_t)m, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return m;
}

// ==========================================================================================
// Multi-GPU plumbing (one process per GPU; the host issues the collectives, e.g. torch.distributed)
// ==========================================================================================
extern "C" int ljmd_ctx_set_stream(ljmd_ctx *c, void *cuda_stream)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_set_stream(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    drop_graphs(c); // graphs were captured on the old stream
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_device_ptrs(ljmd_ctx *c, double **pos0, double **pos1, double **vel, double **acc,
                                    double **obs)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_device_ptrs(): ctx is NULL.");
    if (c->method == LJMD_METHOD_NLIST)
        return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_device_ptrs(): the neighbour-list path keeps its state in cell-sorted order; use ljmd_ctx_download.");
    if (c->method == LJMD_METHOD_ALLPAIRS && c->pw.sorted)
        return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_device_ptrs(): this all-pairs context keeps its state in a spatially sorted order; create it with LJMD_FLAG_NO_SORT or use ljmd_ctx_download.");
    if (pos0) *pos0 = c->rbuf[0];
    if (pos1) *pos1 = c->rbuf[1];
    if (vel) *vel = c->st.vx;
    if (acc) *acc = c->st.ax;
    if (obs) *obs = c->fw.obs;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_set_shard(ljmd_ctx *c, int32_t i_begin, int32_t i_end)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_set_shard(): ctx is NULL.");
    if (c->method != LJMD_METHOD_CELLS || c->R != 1)
        return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_set_shard(): i-range sharding needs the cell-list method and one replica.");
    if (i_begin < 0 || i_end > c->p.n || i_begin >= i_end)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_set_shard(): bad range [%d, %d).", i_begin, i_end);
    drop_graphs(c);
    c->sharded = !(i_begin == 0 && i_end == c->p.n);
    c->shard_begin = i_begin;
    c->shard_end = i_end;
    c->shard_now = 0;
    return LJMD_OK;
}

// Sharded stepping protocol (every rank holds ALL positions, integrates only its own atoms):
//   ljmd_ctx_compute_forces()            forces on the owned atoms at t=0
//   ljmd_ctx_shard_begin(&buf)           owned r(t+dt) -> position buffer `buf`
//   repeat:  [host: all-gather buffer `buf` over the ranks]
//            ljmd_ctx_shard_step(&buf)   force + finalize for the owned atoms from buffer `buf`;
//                                        returns the buffer that now holds the owned r(t+2dt)
// The observables left in the obs buffer are this rank's share without tail corrections.
extern "C" int ljmd_ctx_shard_begin(ljmd_ctx *c, int32_t *buf_to_gather)
{
    clear_error();
    if (!c || !buf_to_gather) return set_error(LJMD_ERR_NULL, "ljmd_ctx_shard_begin(): NULL argument.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(launch_drift(c, 0, 1));
    c->shard_now = 1;
    *buf_to_gather = 1;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_shard_step(ljmd_ctx *c, int32_t *buf_to_gather)
{
    clear_error();
    if (!c || !buf_to_gather) return set_error(LJMD_ERR_NULL, "ljmd_ctx_shard_step(): NULL argument.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    const int now = c->shard_now, prev = 1 - now;
    LJMD_TRY(enqueue_force(c, now));
    LJMD_TRY(launch_finalize(c, 1, prev, now, prev));
    c->shard_now = prev;
    *buf_to_gather = prev;
    return LJMD_OK;
}

// Ends a sharded run: makes position buffer 0 the current one again (the buffer the last force
// call read holds the complete r(t)).  The caller must NOT have gathered the speculative buffer.
extern "C" int ljmd_ctx_shard_end(ljmd_ctx *c)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_shard_end(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    const int cur = 1 - c->shard_now; // the buffer the last force evaluation used
    if (cur != 0)
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->rbuf[0], c->rbuf[1], sizeof(double) * 3 * c->nR, cudaMemcpyDeviceToDevice, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->shard_now = 0;
    return LJMD_OK;
}

extern "C" int ljmd_tail_corrections(const ljmd_params *p, double *tail_epot, double *tail_d_epot, double *tail_dd_epot)
{
    clear_error();
    LJMD_TRY(check_params(p, "ljmd_tail_corrections"));
    if (!tail_epot || !tail_d_epot || !tail_dd_epot) return set_error(LJMD_ERR_NULL, "ljmd_tail_corrections(): NULL output.");
    tail_constants(*p, tail_epot, tail_d_epot, tail_dd_epot);
    return LJMD_OK;
}

extern "C" int ljmd_ctx_method(const ljmd_ctx *c) { return c ? c->method : -1; }
extern "C" int ljmd_ctx_cells_per_edge(const ljmd_ctx *c) { return c ? (c->method == LJMD_METHOD_NLIST ? c->nw.nc : c->cw.nc) : 0; }
extern "C" int64_t ljmd_ctx_launch_count(const ljmd_ctx *c) { return c ? c->launches : 0; }

extern "C" int ljmd_ctx_force_timing(ljmd_ctx *c, int32_t enable, double *ms_total, int64_t *n_evals)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_force_timing(): ctx is NULL.");
    if (ms_total) *ms_total = c->force_ms;
    if (n_evals) *n_evals = c->force_evals;
    c->timing = enable != 0;
    c->force_ms = 0.0;
    c->force_evals = 0;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_pair_counts(ljmd_ctx *c, int64_t *n_candidates, int64_t *n_in_cutoff)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_pair_counts(): ctx is NULL.");
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    if (!c->fw.pair_counts) LJMD_CUDA_CHECK(cudaMalloc(&c->fw.pair_counts, 2 * sizeof(unsigned long long) * c->R));
    LJMD_CUDA_CHECK(cudaMemsetAsync(c->fw.pair_counts, 0, 2 * sizeof(unsigned long long) * c->R, c->stream));
    int rc = (c->method == LJMD_METHOD_NLIST) ? nlist_pair_counts(c, c->fw.pair_counts)
                                              : enqueue_force(c, 0); // positions unchanged: recomputes the same partials, plus counts
    unsigned long long h[2] = {0, 0};
    if (rc == LJMD_OK) {
        cudaError_t e = cudaMemcpyAsync(h, c->fw.pair_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "pair count download failed: %s", cudaGetErrorString(e));
    }
    cudaFree(c->fw.pair_counts);
    c->fw.pair_counts = nullptr;
    if (rc != LJMD_OK) return rc;
    if (c->method == LJMD_METHOD_ALLPAIRS) {
        const int64_t n = c->p.n;
        if (n_candidates) *n_candidates = n * (n - 1) / 2 * c->R;
        // the full-matrix kernel counts ordered pairs
        if (n_in_cutoff) *n_in_cutoff = (int64_t)(c->pw.variant == 0 ? h[1] / 2 : h[1]);
    } else {
        // the cell kernel visits ordered pairs
        if (n_candidates) *n_candidates = (int64_t)(h[0] / 2);
        if (n_in_cutoff) *n_in_cutoff = (int64_t)(h[1] / 2);
    }
    return LJMD_OK;
}

extern "C" int ljmd_bench_chunk_cycles(int32_t ri, int32_t rj, int32_t dense, int32_t warps_per_cta, double *cycles_per_step)
{
    clear_error();
    if (!cycles_per_step) return set_error(LJMD_ERR_NULL, "ljmd_bench_chunk_cycles(): NULL argument.");
    if (ljmd_device_count() == 0) return set_error(LJMD_ERR_CUDA, "ljmd_bench_chunk_cycles(): no CUDA device available.");
    return bench_chunk_cycles(ri, rj, dense, warps_per_cta, cycles_per_step);
}

extern "C" int ljmd_measure_fp64_peak(int32_t device, double *tflops)
{
    clear_error();
    if (!tflops) return set_error(LJMD_ERR_NULL, "ljmd_measure_fp64_peak(): tflops is NULL.");
    if (ljmd_device_count() == 0)
        return set_error(LJMD_ERR_CUDA, "ljmd_measure_fp64_peak(): no CUDA device available.");
    return measure_fp64_peak(device, tflops);
}

extern "C" int ljmd_measure_fp64_peak_sustained(int32_t device, double seconds, double *tflops)
{
    clear_error();
    if (!tflops) return set_error(LJMD_ERR_NULL, "ljmd_measure_fp64_peak_sustained(): tflops is NULL.");
    if (ljmd_device_count() == 0)
        return set_error(LJMD_ERR_CUDA, "ljmd_measure_fp64_peak_sustained(): no CUDA device available.");
    return measure_fp64_peak_sustained(device, seconds, tflops);
}

// ==========================================================================================
// Stateless drop-in pair.  A cached context keeps the device buffers between calls.
// ==========================================================================================
namespace {
std::mutex g_compat_mu;
ljmd_ctx *g_compat = nullptr;

// Page-locking of the CALLER's arrays (large systems only): the drop-in pair is called with the same nine
// allocatable arrays for a whole run (md_simulation_program.f90:216-222 allocates them once), so they are
// registered with the driver at the first call and the H2D / D2H copies then go straight from / to them -- no
// staging memcpy through a pinned bounce buffer (two passes over 75 MB each at N = 2^20: 60 % of the call).
// The registrations are dropped by ljmd_compat_release(), which a host must call BEFORE it frees arrays it has
// passed in; LJMD_COMPAT_PIN=0 turns the feature off.  Small arrays keep the bounce buffer.
struct PinnedRange { const void *p; size_t bytes; };
std::vector<PinnedRange> g_pinned;
const size_t kPinMinBytes = (size_t)1 << 20;

bool compat_pin(const void *p, size_t bytes)
{
    static const bool enabled = !(getenv("LJMD_COMPAT_PIN") && atoi(getenv("LJMD_COMPAT_PIN")) == 0);
    if (!enabled || bytes < kPinMinBytes) return false;
    for (const PinnedRange &r : g_pinned)
        if (r.p == p && r.bytes >= bytes) return true;
    if (g_pinned.size() >= 64) return false;
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError(); // e.g. already registered by the caller with another size: use the bounce buffer
        return false;
    }
    g_pinned.push_back({p, bytes});
    return true;
}

void compat_unpin_all()
{
    for (const PinnedRange &r : g_pinned) cudaHostUnregister(const_cast<void *>(r.p));
    cudaGetLastError();
    g_pinned.clear();
}

int compat_ctx(const ljmd_params *p, const char *who, ljmd_ctx **out)
{
    LJMD_TRY(check_params(p, who));
    if (g_compat && (g_compat->p.n != p->n)) { ljmd_ctx_destroy(g_compat); g_compat = nullptr; }
    if (g_compat && g_compat->method == LJMD_METHOD_CELLS && cells_per_edge_host(*p) != g_compat->cw.nc) {
        ljmd_ctx_destroy(g_compat); g_compat = nullptr;
    }
    if (g_compat && g_compat->method == LJMD_METHOD_NLIST &&
        (nlist_cells_per_edge_host(*p, g_compat->nw.skin) != g_compat->nw.nc || p->rc != g_compat->p.rc ||
         p->density != g_compat->p.density)) {
        ljmd_ctx_destroy(g_compat); g_compat = nullptr;
    }
    if (!g_compat) {
        ljmd_options o = {LJMD_METHOD_AUTO, -1, 1, LJMD_FLAG_NO_SORT}; // host arrays in/out every call: caller's order
        LJMD_TRY(ljmd_ctx_create(p, &o, &g_compat));
    } else if (memcmp(&g_compat->p, p, sizeof(*p)) != 0) {
        g_compat->p = *p;
        refresh_constants(g_compat);
        drop_graphs(g_compat);
        g_compat->nw.compat_keep = false; // the list was built for other parameters
    }
    *out = g_compat;
    return LJMD_OK;
}
} // namespace

extern "C" int ljmd_compat_release(void)
{
    std::lock_guard<std::mutex> lk(g_compat_mu);
    if (g_compat) {
        cudaStreamSynchronize(g_compat->stream);
        compat_unpin_all();
        ljmd_ctx_destroy(g_compat);
        g_compat = nullptr;
    }
    compat_unpin_all();
    return LJMD_OK;
}

extern "C" int ljmd_compute_lj(const ljmd_params *params, const double *rx, const double *ry,
                               const double *rz, double *ax, double *ay, double *az,
                               double *epot, double *d_epot, double *dd_epot)
{
    clear_error();
    std::lock_guard<std::mutex> lk(g_compat_mu);
    LJMD_TRY(check_params(params, "compute_lj_potential_energy"));
    if (!rx || !ry || !rz || !ax || !ay || !az)
        return set_error(LJMD_ERR_NULL, "compute_lj_potential_energy(): state arrays are not allocated.");
    if (!epot || !d_epot || !dd_epot)
        return set_error(LJMD_ERR_NULL, "compute_lj_potential_energy(): output scalar is NULL.");
    ljmd_ctx *c = nullptr;
    LJMD_TRY(compat_ctx(params, "compute_lj_potential_energy", &c));
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(ensure_stage(c));
    const size_t nb = sizeof(double) * c->nR;
    const bool keep_after = c->nw.compat_keep; // a successful force call leaves a consistent sorted state behind
    c->nw.compat_keep = false;
    (void)keep_after;
    const double *in3[3] = {rx, ry, rz};
    double *out3[3] = {ax, ay, az};
    bool direct = true;
    for (int k = 0; k < 3; ++k) direct = compat_pin(in3[k], nb) && direct;
    for (int k = 0; k < 3; ++k) direct = compat_pin(out3[k], nb) && direct;
    if (direct) {
        for (int k = 0; k < 3; ++k)
            LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.rx + (size_t)k * c->nR, in3[k], nb, cudaMemcpyHostToDevice, c->stream));
        const double *dev_acc = c->st.ax;
        if (c->method == LJMD_METHOD_NLIST) {
            LJMD_TRY(nlist_reset_identity(c));
            LJMD_TRY(nlist_enqueue_forces(c));
            LJMD_TRY(nlist_gather_orig(c)); // accelerations back in original order
            dev_acc = c->nw.alt + 6 * c->nR;
        } else {
            LJMD_TRY(enqueue_force(c, 0));
            LJMD_TRY(launch_finalize(c, 0, 0, 0, 0));
        }
        for (int k = 0; k < 3; ++k)
            LJMD_CUDA_CHECK(cudaMemcpyAsync(out3[k], dev_acc + (size_t)k * c->nR, nb, cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_obs, c->fw.obs, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        *epot = c->h_obs[0]; *d_epot = c->h_obs[2]; *dd_epot = c->h_obs[3];
        if (c->method == LJMD_METHOD_NLIST) LJMD_TRY(nlist_check_overflow(c));
        c->nw.compat_keep = c->method == LJMD_METHOD_NLIST;
        return LJMD_OK;
    }
    memcpy(c->h_stage, rx, nb);
    memcpy(c->h_stage + c->nR, ry, nb);
    memcpy(c->h_stage + 2 * c->nR, rz, nb);
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.rx, c->h_stage, 3 * nb, cudaMemcpyHostToDevice, c->stream));
    if (c->method == LJMD_METHOD_NLIST) {
        LJMD_TRY(nlist_reset_identity(c));
        LJMD_TRY(nlist_enqueue_forces(c));
        LJMD_TRY(nlist_gather_orig(c)); // accelerations back in original order
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + 6 * c->nR, c->nw.alt + 6 * c->nR, 3 * nb,
                                        cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + 9 * c->nR, c->fw.obs, sizeof(double) * 4,
                                        cudaMemcpyDeviceToHost, c->stream));
    } else {
        LJMD_TRY(enqueue_force(c, 0));
        LJMD_TRY(launch_finalize(c, 0, 0, 0, 0));
        // accelerations and obs are contiguous: ax ay az | obs
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + 6 * c->nR, c->st.ax, 3 * nb + sizeof(double) * 4,
                                        cudaMemcpyDeviceToHost, c->stream));
    }
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    memcpy(ax, c->h_stage + 6 * c->nR, nb);
    memcpy(ay, c->h_stage + 7 * c->nR, nb);
    memcpy(az, c->h_stage + 8 * c->nR, nb);
    const double *o = c->h_stage + 9 * c->nR;
    *epot = o[0]; *d_epot = o[2]; *dd_epot = o[3];
    if (c->method == LJMD_METHOD_NLIST) LJMD_TRY(nlist_check_overflow(c));
    c->nw.compat_keep = c->method == LJMD_METHOD_NLIST;
    return LJMD_OK;
}

extern "C" int ljmd_verlet_step(const ljmd_params *params, double *rx, double *ry, double *rz,
                                double *vx, double *vy, double *vz, double *ax, double *ay,
                                double *az, double *epot, double *ekin, double *d_epot,
                                double *dd_epot)
{
    clear_error();
    std::lock_guard<std::mutex> lk(g_compat_mu);
    if (params && params->n <= 0) return set_error(LJMD_ERR_N, "verlet_step(): params%%n must be > 0.");
    if (!rx || !ry || !rz || !vx || !vy || !vz || !ax || !ay || !az)
        return set_error(LJMD_ERR_NULL, "verlet_step(): state arrays are not allocated.");
    if (!epot || !ekin || !d_epot || !dd_epot)
        return set_error(LJMD_ERR_NULL, "verlet_step(): output scalar is NULL.");
    ljmd_ctx *c = nullptr;
    LJMD_TRY(compat_ctx(params, "verlet_step", &c));
    DeviceGuard g;
    LJMD_TRY(g.enter(c->device));
    LJMD_TRY(ensure_stage(c));
    const size_t nb = sizeof(double) * c->nR;
    double *arr[9] = {rx, ry, rz, vx, vy, vz, ax, ay, az};
    bool direct = true; // large arrays: copy straight from / to the caller's (now page-locked) arrays
    for (int k = 0; k < 9; ++k) direct = compat_pin(arr[k], nb) && direct;
    // LJMD_COMPAT_KEEP_LIST=0: re-sort and rebuild the neighbour list in every call (the behaviour of round 1)
    const char *keep_env = getenv("LJMD_COMPAT_KEEP_LIST");
    const bool keep_enabled = !(keep_env && atoi(keep_env) == 0);
    const bool keep = c->method == LJMD_METHOD_NLIST && keep_enabled && c->nw.compat_keep;
    c->nw.compat_keep = false; // until this call has gone through
    if (direct) {
        const double *dev_src = c->state_block;
        // LJMD_COMPAT_OVERLAP=0: download r(t+dt) after the step with v and a (the behaviour before the overlap)
        static const bool overlap = !(getenv("LJMD_COMPAT_OVERLAP") && atoi(getenv("LJMD_COMPAT_OVERLAP")) == 0);
        int first_out = 0; // first array still to be copied to the caller on the main stream
        double *h2d_dst = keep ? c->nw.alt : c->state_block;
        for (int k = 0; k < 9; ++k)
            LJMD_CUDA_CHECK(cudaMemcpyAsync(h2d_dst + (size_t)k * c->nR, arr[k], nb, cudaMemcpyHostToDevice, c->stream));
        if (c->method == LJMD_METHOD_NLIST) {
            if (keep) LJMD_TRY(nlist_load_from_orig(c));
            else LJMD_TRY(nlist_reset_identity(c));
            LJMD_TRY(nlist_enqueue_drift(c));
            if (overlap) { // r(t+dt) is final now: its download overlaps the force / kick kernel
                LJMD_TRY(nlist_early_positions_d2h(c, arr));
                first_out = 3;
            }
            int rc = nlist_enqueue_step(c);
            if (rc == LJMD_OK) rc = first_out ? nlist_gather_orig_va(c) // v, a back in original order
                                              : nlist_gather_orig(c);   // r(t+dt), v, a back in original order
            if (rc != LJMD_OK) { // nothing may still be writing into the caller's arrays when the error is returned
                if (first_out) cudaStreamSynchronize(c->nw.side_stream);
                return rc;
            }
            dev_src = c->nw.alt;
        } else {
            LJMD_TRY(launch_drift(c, 0, 1));
            LJMD_TRY(enqueue_step(c, 0, 2));
        }
        for (int k = first_out; k < 9; ++k)
            LJMD_CUDA_CHECK(cudaMemcpyAsync(arr[k], dev_src + (size_t)k * c->nR, nb, cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_obs, c->fw.obs, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream));
        if (first_out) LJMD_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->nw.ev_r_copied, 0));
        LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (c->method == LJMD_METHOD_NLIST) LJMD_TRY(nlist_check_overflow(c));
        c->nw.compat_keep = true;
        *epot = c->h_obs[0]; *ekin = c->h_obs[1]; *d_epot = c->h_obs[2]; *dd_epot = c->h_obs[3];
        return LJMD_OK;
    }
    for (int k = 0; k < 9; ++k) memcpy(c->h_stage + (size_t)k * c->nR, arr[k], nb);
    // Upload, drift, force, finalize and download replayed as ONE graph launch instead of two copies
    // and three kernel launches (bitwise equal to the separate launches, tests/test_gpu_parity.py;
    // 97 -> 87 us per call at N = 4000).  LJMD_COMPAT_GRAPH=0 selects the separate launches.
    static const bool use_graph = !(getenv("LJMD_COMPAT_GRAPH") && atoi(getenv("LJMD_COMPAT_GRAPH")) == 0);
    if (use_graph && c->method != LJMD_METHOD_NLIST) {
        if (!c->graph_compat) {
            cudaGraph_t g = nullptr;
            LJMD_CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            c->capturing = true;
            cudaError_t e = cudaMemcpyAsync(c->state_block, c->h_stage, 9 * nb, cudaMemcpyHostToDevice, c->stream);
            int rc = (e == cudaSuccess) ? launch_drift(c, 0, 1) : LJMD_OK;
            if (e == cudaSuccess && rc == LJMD_OK) rc = enqueue_step(c, 0, 2);
            if (e == cudaSuccess && rc == LJMD_OK)
                e = cudaMemcpyAsync(c->h_stage, c->state_block, 9 * nb + sizeof(double) * 4, cudaMemcpyDeviceToHost,
                                    c->stream);
            c->capturing = false;
            const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
            if (rc != LJMD_OK || e != cudaSuccess || e2 != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                if (rc != LJMD_OK) return rc;
                LJMD_CUDA_CHECK(e);
                LJMD_CUDA_CHECK(e2);
            }
            const cudaError_t e3 = cudaGraphInstantiate(&c->graph_compat, g, 0);
            cudaGraphDestroy(g);
            LJMD_CUDA_CHECK(e3);
        }
        LJMD_CUDA_CHECK(cudaGraphLaunch(c->graph_compat, c->stream));
        c->launches += 1 + kernels_per_step(c);
        LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 9; ++k) memcpy(arr[k], c->h_stage + (size_t)k * c->nR, nb);
        const double *og = c->h_stage + 9 * c->nR;
        *epot = og[0]; *ekin = og[1]; *d_epot = og[2]; *dd_epot = og[3];
        return LJMD_OK;
    }
    LJMD_CUDA_CHECK(cudaMemcpyAsync(keep ? c->nw.alt : c->state_block, c->h_stage, 9 * nb, cudaMemcpyHostToDevice, c->stream));
    if (c->method == LJMD_METHOD_NLIST) {
        if (keep) LJMD_TRY(nlist_load_from_orig(c));
        else LJMD_TRY(nlist_reset_identity(c));
        LJMD_TRY(nlist_enqueue_drift(c));
        LJMD_TRY(nlist_enqueue_step(c));
        LJMD_TRY(nlist_gather_orig(c)); // r(t+dt), v, a back in original order
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage, c->nw.alt, 9 * nb, cudaMemcpyDeviceToHost, c->stream));
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage + 9 * c->nR, c->fw.obs, sizeof(double) * 4,
                                        cudaMemcpyDeviceToHost, c->stream));
    } else {
        LJMD_TRY(launch_drift(c, 0, 1));
        LJMD_TRY(enqueue_step(c, 0, 2)); // mode 2: the new positions are copied back into rbuf[0]
        LJMD_CUDA_CHECK(cudaMemcpyAsync(c->h_stage, c->state_block, 9 * nb + sizeof(double) * 4,
                                        cudaMemcpyDeviceToHost, c->stream));
    }
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->method == LJMD_METHOD_NLIST) LJMD_TRY(nlist_check_overflow(c));
    c->nw.compat_keep = true;
    for (int k = 0; k < 9; ++k) memcpy(arr[k], c->h_stage + (size_t)k * c->nR, nb);
    const double *o = c->h_stage + 9 * c->nR;
    *epot = o[0]; *ekin = o[1]; *d_epot = o[2]; *dd_epot = o[3];
    return LJMD_OK;
}
