// ljmd_internal.h -- internal declarations shared by the .cu files of libljmd.so.
// Not part of the public ABI (include/ljmd.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ljmd.h"

namespace ljmd {

// ------------------------------------------------------------------------------------------
// Error plumbing: every failure records a thread-local message (ljmd_last_error()).
// ------------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void clear_error();

#define LJMD_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t err__ = (expr);                                                        \
        if (err__ != cudaSuccess)                                                          \
            return ::ljmd::set_error(LJMD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                                     cudaGetErrorString(err__), __FILE__, __LINE__);       \
    } while (0)

#define LJMD_TRY(expr)                  \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != LJMD_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------
// Device-side control block, read by the finalize kernel (fixed address => CUDA graphs can be
// replayed while the host changes where the per-step observables go).
// ------------------------------------------------------------------------------------------
struct StepCtl {
    long long step;       // steps finished since the host last reset it
    long long series_cap; // capacity of `series` in steps (0 = not recording)
    double *series;       // [cap][R][4]
};

// Constants every kernel needs, passed by value.
struct SimConst {
    int n;            // atoms per replica
    int R;            // replicas
    double L, invL;   // box
    double rc2;       // cutoff^2
    double dt, dt_half, dt_square_half;
    double tail_u, tail_du, tail_ddu; // lj_potential_energy.f90:205-223, host-computed
};

// State arrays, each R*n doubles.
struct StatePtrs {
    double *rx, *ry, *rz, *vx, *vy, *vz, *ax, *ay, *az;
    double *rux, *ruy, *ruz; // nullptr when unwrapped tracking is off
};

// Work buffers of the all-pairs path.
struct PairWork {
    int block;       // threads per CTA of the pair kernel
    int ipt;         // i-atoms per thread
    int n_itiles;    // CTAs along i
    int n_splits;    // CTAs along j (each handles a contiguous j range)
    int split_len;   // atoms per j split
    double *fpart;   // [R][n_splits][3][n] partial forces (without the factor 24)
    double *spart;   // [R][n_itiles*n_splits][3] partial (u, du, ddu) over ORDERED pairs
    int variant;     // 0 = full-matrix kernel, 1 = symmetric-tile kernel
};

struct FinalizeWork {
    int block;          // threads per CTA
    int nblk;           // CTAs per replica
    double *kpart;      // [R][nblk][3] partial sums of vx^2, vy^2, vz^2
    unsigned *counter;  // [R] last-block tickets
    double *obs;        // [R][4] observables of the last step: epot, ekin, d_epot, dd_epot
    StepCtl *ctl;       // device control block
    unsigned long long *pair_counts; // [2] candidates, in-cutoff (debug/bench, may be null)
};

// Cell-list work buffers (lj_cells.cu).
struct CellWork {
    int nc;             // cells per edge (0 = unused)
    int ncells;         // nc^3
    double scale;       // nc / L  (x*scale -> cell coordinate)
    int *cell_of_atom;  // [R][n]
    int *cell_start;    // [R][ncells+1]
    int *cell_fill;     // [R][ncells] scratch
    int *order;         // [R][n] sorted position -> original index
    double *sx, *sy, *sz; // [R][n] positions in cell-sorted order
    double *spart;      // [R][nblk_force][3]
    int nblk_force;
    int block;
    void *scan_tmp;     // scratch for the block scan
    unsigned long long *pair_counts; // [R][2]
    double *fsorted;    // [R][3][n] forces in sorted order (without factor 24)
};

} // namespace ljmd

// The opaque context of the public ABI.
struct ljmd_ctx {
    ljmd_params p;
    ljmd_options o;
    int device;
    int method; // resolved
    int R;
    size_t nR;  // n*R
    cudaStream_t stream;
    ljmd::SimConst sc;
    ljmd::StatePtrs st;
    double *state_block; // one allocation backing st.* (9 or 12 arrays)
    ljmd::PairWork pw;
    ljmd::FinalizeWork fw;
    ljmd::CellWork cw;
    double *h_obs; // pinned [R][4] (+ scratch)
    double *h_stage; // pinned staging for host<->device state copies (9*nR doubles)
    // graphs
    cudaGraphExec_t graph_chunk; // G steps
    int graph_chunk_steps;
    cudaGraphExec_t graph_one;   // 1 step
    // series buffer
    double *d_series;
    long long series_cap;
    // accounting
    int64_t launches;
    bool timing;
    double force_ms;
    int64_t force_evals;
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_s0, ev_s1; // bracket the last ljmd_ctx_step call on the stream
    double last_step_ms;
};

namespace ljmd {

// ---- verlet.cu ---------------------------------------------------------------------------
// drift + wrap + first half-kick (+ unwrapped update), verlet.f90:58-74
int launch_verlet_pre(ljmd_ctx *c);
// a = 24*sum(partials); optional second half-kick + E_kin (verlet.f90:86-95); scalar reduction
int launch_finalize(ljmd_ctx *c, bool kick);
int launch_ekin_driver(ljmd_ctx *c, double *d_out);
int measure_fp64_peak(int device, double *tflops);

// ---- lj_allpairs.cu ----------------------------------------------------------------------
int pairwork_plan(ljmd_ctx *c);
int launch_pairs(ljmd_ctx *c);

// ---- lj_cells.cu -------------------------------------------------------------------------
int cellwork_plan(ljmd_ctx *c);
void cellwork_free(ljmd_ctx *c);
int launch_cells_force(ljmd_ctx *c);
int cells_per_edge_host(const ljmd_params &p);

} // namespace ljmd
