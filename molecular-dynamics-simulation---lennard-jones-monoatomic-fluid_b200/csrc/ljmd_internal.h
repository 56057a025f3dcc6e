// ljmd_internal.h -- internal declarations shared by the .cu files of libljmd.so.
// Not part of the public ABI (include/ljmd.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ljmd.h"
#include "ljmd_debug.h"

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
    int variant;     // 0 = full-matrix kernel, 1 = symmetric-tile kernel (default)
    // ---- variant 0 ----
    int block;       // threads per CTA
    int ipt;         // i-atoms per thread
    int n_itiles;    // CTAs along i
    int split_len;   // atoms per j split
    // ---- variant 1 ----
    int ri, warps;        // i-atoms per lane, warps per CTA (row = ri*32*warps atoms)
    int rj;               // rotation offsets per step (pairs per lane-step = ri*rj)
    int n_rows;           // rows of i-atoms
    int chunks_per_item;  // (unused by the symmetric kernel; kept for the finalize argument block)
    int *row_splits;      // device [n_rows]: number of work items (= i-side partials) of each row
    bool dense;           // large in-cutoff fraction: interaction body always executed
    bool indep;           // independent-warp kernel (rows of ri*32 atoms, fat CTAs of cta_warps)
    int cta_warps;        // warps per CTA of the independent-warp kernel
    int n_ctas;           // CTAs per replica of the launch
    int n_spart;          // scalar partials per replica
    long long *dbg;       // optional per-warp timestamps of the independent-warp kernel (tuning)
    int4 *items;          // device work-item table
    double *fJ;           // [R][n_rows][3][n] j-side partial forces
    // ---- variant 1, spatially sorted flavour (sparse cutoff, large N; lj_pairs_sorted_kernel) ----
    bool sorted;          // the resident state is kept in a spatially sorted order (perm below)
    int *perm;            // HOST: perm[slot] = original atom index (nullptr = identity)
    long long steps_since_sort;
    double *cbox;         // device [nchunks][6]: bounding box (lo xyz, hi xyz) of every 32-atom chunk
    unsigned char *near;  // device [nchunks][n_rows]: 1 if the (row, chunk) tile wrote a j-side partial this step
    unsigned *work;       // device [2]: dynamic work-item counter, finished-warp counter
    // ---- both ----
    int n_splits;    // variant 0: j splits; variant 1: max splits of any row
    int n_items;     // CTAs per replica (= scalar partials per replica)
    double *fpart;   // variant 0: [R][n_splits][3][n]; variant 1: i-side partials, same shape
    double *spart;   // [R][n_items][3] scalar partials
};

struct FinalizeWork {
    int block;          // threads per CTA
    int groups;         // thread groups sharing one atom's partial sum (1 or 8)
    int nblk;           // CTAs per replica
    double *kpart;      // [R][nblk][6] per-CTA partial sums: vx^2, vy^2, vz^2 and a slice of the force kernel's scalar partials
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
    int *cell_count;    // [R][ncells] histogram
    int G, nzg, nunits; // cells per z-group, z-groups per column, warp work units
    int *order;         // [R][n] sorted position -> original index
    double *sx, *sy, *sz; // [R][n] positions in cell-sorted order
    double *spart;      // [R][nblk_force][3]
    int nblk_force;
    int kernels_per_force; // launches per force evaluation (binning + sort + force)
    int block;
    void *scan_tmp;     // scratch for the block scan
    unsigned long long *pair_counts; // [R][2]
    double *fsorted;    // [R][3][n] forces in sorted order (without factor 24)
    int *slot_of_atom;  // [R][n] inverse of `order`
    double pair_weight; // 0.5 if the force kernel visits ordered pairs, 1.0 if unordered
};

// Device control block of the neighbour-list path (lj_nlist.cu).  The fields up to `rebuilds`
// are reset together when the host puts the state back into original order.
struct NlCtl {
    int need_rebuild;      // set by the drift / step epilogue (max displacement > skin/2) or the host
    int cur;               // index of the position buffer that holds r(t); flips every step
    int overflow;          // largest neighbour count seen if it exceeded the capacity, else 0
    int unwrapped;         // the last build saw positions outside [0, L]: minimum image everywhere
    unsigned ticket_step, ticket_build, ticket_drift; // last-CTA tickets
    unsigned pad_;
    unsigned long long maxd2_bits; // max squared displacement since the last build (double bits)
    long long rebuilds;    // number of list builds so far
    // ---- slab sharding (lj_nlist.cu)
    int halted;            // a rank reported a skin violation: step kernels are no-ops until the host rebuilds
    int steps_left;        // step kernels still allowed to run in the current ljmd_ctx_slab_run call
    int steps_done;        // steps completed in the current ljmd_ctx_slab_run call
    int seq;               // sharded steps completed since ljmd_ctx_slab_attach
    int pending_flag;      // flag value of the last completed step, published by the next launch
};

// Neighbour-list work buffers (lj_nlist.cu).  The resident state is kept in cell-sorted order.
struct NlistWork {
    int nc, ncells;      // cells per edge (edge >= rc + skin), nc^3
    double scale;        // nc / L
    double skin, rl2, half_skin2;
    int maxnb;           // list capacity per atom
    size_t stride;       // slots per entry group (n rounded up to 32); see nl_entry() in lj_nlist.cu
    int nblk, narr;
    int tpa;             // lanes per atom of the step kernel (1 or 4)
    int pipe;            // software-pipelined gathers (small systems)
    int cur_host;        // host mirror of NlCtl::cur
    bool compat_keep;    // drop-in calls: the order and list left by the previous call can be reused
    long long rebuilds_host, rebuilds_counted;
    int *cell_of_slot, *cell_count, *cell_start, *cell_cursor, *tile_sum;
    unsigned long long *keys;
    int *orig, *orig_alt, *slot_of_orig;
    unsigned char *interior;
    int *nb_cnt;
    int *nbr;              // the list (global slot indices, groups of four)
    int nzg, nunits, cap_stage;
    size_t smem_bytes;
    double *r0, *alt, *spart;
    double *mom;         // [3] momentum of the atoms handled by the last step launch
    double4 *p4;         // [2][n] packed positions (x, y, z, 0) of the two position buffers
    NlCtl *ctl;
    cudaStream_t side_stream; // captures the IF-node bodies of the step graph; carries the early position download of drop-in calls
    double *compat_r;         // [3][n] r(t+dt) in original order, gathered right after the drift of a drop-in call (lazily allocated)
    cudaEvent_t ev_r_ready, ev_r_copied; // drop-in calls: compat_r written (main stream) / copied to the caller (side stream)
    // slab sharding over the GPUs of a node
    int sh_on, rank, world, xlo, xhi, own_blocks;
    double4 *peer_p4[LJMD_MAX_RANKS];
    int *peer_flags[LJMD_MAX_RANKS];
    void *ipc_p4[LJMD_MAX_RANKS], *ipc_flags[LJMD_MAX_RANKS];
    int *flags;
    int *sl_rng;         // device [16]: slot pieces of a slab-local rebuild (see NlArgs::sl_rng)
    int sl_cand_cap, sl_halo_cap; // launch sizes: candidate slots of a local rebuild, atoms of two cell layers
    int *dbg_exec;       // debug counters (LJMD_SLAB_DEBUG)
    unsigned long long *dbg_time; // per-step device timestamps (LJMD_SLAB_TIMELINE)
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
    double *state_block; // one allocation backing st.*: 9 arrays | obs | alt positions | (ru)
    double *rbuf[2];     // position buffers (x array; y, z follow at +nR, +2nR).  rbuf[0] is the
                         // state block's rx..rz and always holds the current positions outside
                         // of ljmd_ctx_step; rbuf[1] is the ping-pong partner.
    bool capturing;      // stream capture in progress (launch counters are not advanced)
    cudaStream_t own_stream; // the stream created with the context (stream may be the caller's)
    bool sharded;        // multi-GPU i-range sharding: this context integrates [shard_begin, shard_end)
    int shard_begin, shard_end;
    int shard_now;       // sharded stepping: index of the position buffer the next force call reads
    ljmd::PairWork pw;
    ljmd::FinalizeWork fw;
    ljmd::CellWork cw;
    ljmd::NlistWork nw;
    double *h_obs; // pinned [R][4] (+ scratch)
    double *h_stage; // pinned staging for host<->device state copies (9*nR doubles)
    // graphs
    cudaGraphExec_t graph_chunk; // G steps
    cudaGraphExec_t graph_compat; // opt-in: one ljmd_verlet_step host-array call (upload, step, download)
    cudaGraphExec_t graph_tail[16]; // graph_tail[k]: 2k steps (the even remainder of a run after the 32-step chunks)
    cudaGraphExec_t graph_slab[33]; // graph_slab[k]: k sharded steps (ljmd_ctx_slab_run with max_steps = k <= 32)
    int graph_chunk_steps;
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
// rbuf[to] = wrap(drift(rbuf[from], v, a)), verlet.f90:58-63
int launch_drift(ljmd_ctx *c, int from, int to);
// a = 24*sum(partials) (+ both half kicks, E_kin, ru, next positions); scalar reduction.
// mode 0: forces only; 1: full step, rbuf[out] = positions of the next step; 2: full step,
// rbuf[out] = copy of rbuf[now].
int launch_finalize(ljmd_ctx *c, int mode, int prev, int now, int out);
int launch_ekin_driver(ljmd_ctx *c, double *d_out);
int measure_fp64_peak(int device, double *tflops);
int measure_fp64_peak_sustained(int device, double seconds, double *tflops);

// ---- lj_allpairs.cu ----------------------------------------------------------------------
int pairwork_plan(ljmd_ctx *c);
int launch_pairs(ljmd_ctx *c, int buf); // positions from rbuf[buf]
int bench_chunk_cycles(int ri, int rj, int dense, int warps_per_cta, double *cycles_per_step);

// ---- lj_cells.cu -------------------------------------------------------------------------
int cellwork_plan(ljmd_ctx *c);
void cellwork_free(ljmd_ctx *c);
int launch_cells_force(ljmd_ctx *c, int buf);
int cells_per_edge_host(const ljmd_params &p);

// ---- lj_nlist.cu -------------------------------------------------------------------------
int nlist_plan(ljmd_ctx *c);
void nlist_free(ljmd_ctx *c);
int nlist_cells_per_edge_host(const ljmd_params &p, double skin);
double nlist_default_skin();
int nlist_request_rebuild(ljmd_ctx *c);
int nlist_enqueue_rebuild(ljmd_ctx *c, int eval_sel);
int nlist_enqueue_forces(ljmd_ctx *c);      // rebuild-if-needed + forces/observables at r(t)
int nlist_enqueue_drift(ljmd_ctx *c);       // first step of a run
int nlist_enqueue_step(ljmd_ctx *c);        // rebuild-if-needed + fused step kernel
int nlist_enqueue_step_kernel(ljmd_ctx *c); // the fused step kernel alone
int nlist_capture_graph(ljmd_ctx *c, int nsteps, cudaGraphExec_t *out);
int nlist_gather_orig(ljmd_ctx *c);         // state in original order -> nw.alt [r v a (ru)]
int nlist_early_positions_d2h(ljmd_ctx *c, double *const host_r[3]); // drop-in calls: r(t+dt) to the caller while the step kernel runs
int nlist_gather_orig_va(ljmd_ctx *c);      // v, a in original order -> nw.alt [. v a]
int nlist_canonicalize(ljmd_ctx *c);
int nlist_reset_identity(ljmd_ctx *c);
int nlist_load_from_orig(ljmd_ctx *c);      // w.alt = [r | v | a] in original order -> resident order (drop-in calls)
int nlist_check_overflow(ljmd_ctx *c);
int nlist_pair_counts(ljmd_ctx *c, unsigned long long *d_counts);
int nlist_download_cells(ljmd_ctx *c, int32_t *cell_of_atom, int32_t *cell_start, int32_t *order);

} // namespace ljmd
