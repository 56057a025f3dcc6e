/*
 * ljmd.h -- C ABI of the B200-native Lennard-Jones MD hot path (libljmd.so).
 *
 * Drop-in boundary for the reference's two hot-path module procedures
 * (Ledicia/Molecular-Dynamics-Simulation---Lennard-Jones-monoatomic-fluid; citations are
 * file:line relative to the reference's scripts/ directory):
 *
 *   compute_lj_potential_energy(params, state, epot, d_epot, dd_epot)
 *                                              physics/lj_potential_energy.f90:46-52
 *   verlet_step(params, state, epot, ekin, d_epot, dd_epot)
 *                                              physics/verlet.f90:41-48
 *
 * The reference has no FFI layer (plain Fortran module procedures).  INTEGRATION.md shows the
 * ISO_C_BINDING replacement modules that keep those names/dummy lists and call the functions
 * below.  All reals are IEEE binary64, all integers int32 (base/define_precision.f90:14,17).
 *
 * Conventions
 *   - Every function returns an int status: LJMD_OK (0) or an LJMD_ERR_* code.  The reference
 *     `stop`s with a message instead; ljmd_last_error() returns the equivalent text so a shim
 *     can reproduce that behaviour.
 *   - The library never retains host pointers after a call returns and never frees them.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     LJMD_ERR_CUDA.
 *   - Host-pointer functions are synchronous.  One host thread per context.
 */
#ifndef LJMD_H
#define LJMD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LJMD_VERSION 100

/* Mirror of type(sim_params), base/md_types.f90:27-50, in gfortran's de-facto layout:
 * offsets 0,4,8,16,...,88; sizeof == 96. */
typedef struct ljmd_params {
    int32_t n;              /* number of particles                      md_types.f90:29 */
    int32_t num_cells;      /* FCC cells per edge                       md_types.f90:30 */
    double box_length;      /* L                                        md_types.f90:33 */
    double inv_box_length;  /* 1/L                                      md_types.f90:34 */
    double volume;          /* L**3                                     md_types.f90:35 */
    double density;         /* n / volume                               md_types.f90:36 */
    double dt;              /* time step                                md_types.f90:39 */
    double dt_half;         /* 0.5*dt                                   md_types.f90:40 */
    double dt_square_half;  /* 0.5*dt*dt                                md_types.f90:41 */
    double rc;              /* cutoff radius                            md_types.f90:44 */
    double rc_square;       /* rc*rc                                    md_types.f90:45 */
    double length_scale;    /* legacy, unused on the hot path           md_types.f90:48 */
    double inv_length_scale;/*                                          md_types.f90:49 */
} ljmd_params;

/* Status codes.  1..5 mirror the `stop` guards of lj_potential_energy.f90:77-81; 6 mirrors the
 * allocated() guards (:82, verlet.f90:52); 7,8 mirror md_types.f90:152,161. */
enum {
    LJMD_OK = 0,
    LJMD_ERR_N = 1,            /* params%n must be > 0 */
    LJMD_ERR_BOX = 2,          /* params%box_length must be > 0 */
    LJMD_ERR_VOLUME = 3,       /* params%volume must be > 0 */
    LJMD_ERR_RC = 4,           /* params%rc must be > 0 */
    LJMD_ERR_RC_SQUARE = 5,    /* params%rc_square must be > 0 */
    LJMD_ERR_NULL = 6,         /* a required array/pointer is NULL ("not allocated") */
    LJMD_ERR_RC_HALF_BOX = 7,  /* rc must be < L/2 (minimum image) */
    LJMD_ERR_DT = 8,           /* dt must be > 0 */
    LJMD_ERR_ARG = 9,          /* invalid option / argument */
    LJMD_ERR_UNSUPPORTED = 10, /* e.g. cell list requested with fewer than 3 cells per edge */
    LJMD_ERR_CUDA = 100,       /* CUDA runtime/driver error (no device, launch failure, ...) */
    LJMD_ERR_NOMEM = 101
};

/* Force method of a context. */
enum {
    LJMD_METHOD_AUTO = 0,     /* all-pairs unless floor(L/(rc+skin)) >= 4 and N >= 8192 (then NLIST) */
    LJMD_METHOD_ALLPAIRS = 1, /* O(N^2), the reference's algorithm */
    LJMD_METHOD_CELLS = 2,    /* counting-sort cell list + 27-cell stencil every step */
    LJMD_METHOD_NLIST = 3     /* counting-sort cells + Verlet neighbour list with skin, one fused
                                 kernel per step, state kept in cell-sorted order on the device
                                 (CELLS and NLIST have no reference counterpart: the reference is
                                 all-pairs only, lj_potential_energy.f90:109-183) */
};

/* Context flags. */
enum {
    LJMD_FLAG_TRACK_UNWRAPPED = 1, /* keep ru(t) on the device, md_simulation_program.f90:339-353 */
    LJMD_FLAG_NO_SORT = 2          /* ALLPAIRS: keep the caller's atom order on the device.  By default a large
                                      sparse-cutoff system (N >= 16384, cutoff sphere < 15 % of the box) is held
                                      in a spatially sorted order -- upload / download translate, results are
                                      order independent up to summation order */
};

typedef struct ljmd_options {
    int32_t method;    /* LJMD_METHOD_* */
    int32_t device;    /* CUDA device ordinal; -1 = the calling thread's current device */
    int32_t nreplicas; /* >= 1 independent systems sharing `params`; arrays are [nreplicas][n] */
    int32_t flags;     /* LJMD_FLAG_* */
} ljmd_options;

typedef struct ljmd_ctx ljmd_ctx;

/* ----------------------------------------------------------------------------------------
 * Library / diagnostics
 * -------------------------------------------------------------------------------------- */
int ljmd_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char *ljmd_last_error(void);
/* Number of CUDA devices visible (0 if none / no driver). */
int ljmd_device_count(void);

/* ----------------------------------------------------------------------------------------
 * Parameters: init_params + compute_derived_params, base/md_types.f90:105-169.
 * Returns LJMD_ERR_BOX / _RC / _RC_HALF_BOX / _DT where the reference stops (:143,150,153,161).
 * -------------------------------------------------------------------------------------- */
int ljmd_params_init(ljmd_params *p, int32_t n, double box_length, double dt, double rc,
                     int32_t num_cells);

/* ----------------------------------------------------------------------------------------
 * Stateless drop-in pair (host pointers, length params->n each; copies in and out every call).
 * -------------------------------------------------------------------------------------- */

/* Replaces compute_lj_potential_energy, physics/lj_potential_energy.f90:46-225.
 * in: rx,ry,rz.  out: ax,ay,az (overwritten), epot, d_epot, dd_epot (tail corrections included,
 * :205-223). */
int ljmd_compute_lj(const ljmd_params *params, const double *rx, const double *ry,
                    const double *rz, double *ax, double *ay, double *az, double *epot,
                    double *d_epot, double *dd_epot);

/* Replaces verlet_step, physics/verlet.f90:41-97.  All nine arrays are updated in place. */
int ljmd_verlet_step(const ljmd_params *params, double *rx, double *ry, double *rz, double *vx,
                     double *vy, double *vz, double *ax, double *ay, double *az, double *epot,
                     double *ekin, double *d_epot, double *dd_epot);

/* Release the cached device workspace the two functions above keep between calls. */
int ljmd_compat_release(void);

/* ----------------------------------------------------------------------------------------
 * Device-resident API (state stays in HBM between calls).
 * -------------------------------------------------------------------------------------- */
int ljmd_ctx_create(const ljmd_params *params, const ljmd_options *opts, ljmd_ctx **out);
int ljmd_ctx_destroy(ljmd_ctx *ctx);

/* Host -> device.  Any pointer may be NULL (that array is left unchanged).  Each array holds
 * nreplicas*n doubles.  Uploading positions also resets ru := r when unwrapped tracking is on
 * (md_simulation_program.f90:226-231). */
int ljmd_ctx_upload(ljmd_ctx *ctx, const double *rx, const double *ry, const double *rz,
                    const double *vx, const double *vy, const double *vz, const double *ax,
                    const double *ay, const double *az);
/* Device -> host.  Any pointer may be NULL. */
int ljmd_ctx_download(ljmd_ctx *ctx, double *rx, double *ry, double *rz, double *vx, double *vy,
                      double *vz, double *ax, double *ay, double *az);
/* Unwrapped coordinates ru (needs LJMD_FLAG_TRACK_UNWRAPPED). */
int ljmd_ctx_download_unwrapped(ljmd_ctx *ctx, double *rux, double *ruy, double *ruz);

/* compute_lj_potential_energy on the resident state (the t=0 call,
 * md_simulation_program.f90:236).  obs3 = {epot, d_epot, dd_epot} per replica (3*nreplicas). */
int ljmd_ctx_compute_forces(ljmd_ctx *ctx, double *obs3);

/* nsteps x verlet_step (+ the unwrapped update if tracked) with no host round-trip.
 * obs4 = {epot, ekin, d_epot, dd_epot} of the LAST step, per replica (4*nreplicas); may be NULL.
 */
int ljmd_ctx_step(ljmd_ctx *ctx, int32_t nsteps, double *obs4);

/* As ljmd_ctx_step, but records the 4 observables of every step: series[nsteps][nreplicas][4]. */
int ljmd_ctx_step_series(ljmd_ctx *ctx, int32_t nsteps, double *series);

/* Per-step recorder for stepping calls other than ljmd_ctx_step_series (e.g. sharded runs):
 * arm with a capacity in steps (0 disarms); steps are counted from the arming call. */
int ljmd_ctx_series_arm(ljmd_ctx *ctx, int32_t cap);
int ljmd_ctx_series_fetch(ljmd_ctx *ctx, int32_t nsteps, double *series);

/* Driver-side kinetic energy 0.5*sum(vx^2+vy^2+vz^2) (md_simulation_program.f90:238-240),
 * per replica. */
int ljmd_ctx_kinetic_energy(ljmd_ctx *ctx, double *ekin);

/* ----------------------------------------------------------------------------------------
 * Multi-GPU plumbing.  One process per GPU; the HOST issues the collectives (torch.distributed /
 * NCCL) on the same stream.  No reference counterpart (the reference is a serial program).
 * -------------------------------------------------------------------------------------- */
/* Run every kernel / copy of this context on the caller's CUDA stream (cudaStream_t); NULL
 * restores the context's own stream. */
int ljmd_ctx_set_stream(ljmd_ctx *ctx, void *cuda_stream);
/* Device addresses: the two position buffers (x array of nreplicas*n doubles, y and z follow),
 * velocities, accelerations (same layout) and the observables [nreplicas][4].  Any may be NULL. */
int ljmd_ctx_device_ptrs(ljmd_ctx *ctx, double **pos0, double **pos1, double **vel, double **acc,
                         double **obs);
/* i-particle-range sharding of a large system (cell-list method): this context integrates atoms
 * [i_begin, i_end) and needs all positions from the other ranks each step. */
int ljmd_ctx_set_shard(ljmd_ctx *ctx, int32_t i_begin, int32_t i_end);
/* Sharded stepping: shard_begin writes the owned r(t+dt) into position buffer *buf; the host
 * all-gathers that buffer; shard_step evaluates forces from it, integrates the owned atoms and
 * returns in *buf the buffer that now holds the owned r(t+2dt) (to be gathered next).  The
 * observables buffer then holds this rank's share (no tail corrections). */
int ljmd_ctx_shard_begin(ljmd_ctx *ctx, int32_t *buf_to_gather);
int ljmd_ctx_shard_step(ljmd_ctx *ctx, int32_t *buf_to_gather);
int ljmd_ctx_shard_end(ljmd_ctx *ctx);
/* Tail corrections of lj_potential_energy.f90:205-223 (added once after a cross-rank reduction). */
int ljmd_tail_corrections(const ljmd_params *params, double *tail_epot, double *tail_d_epot,
                          double *tail_dd_epot);
/* ---- Slab sharding of the neighbour-list path (LJMD_METHOD_NLIST) over the GPUs of one node.
 * The cell-sorted slot order makes contiguous slot ranges spatial slabs: rank r owns the cell
 * layers cx in [r*nc/world, (r+1)*nc/world).  Every step kernel writes the new positions of its
 * boundary layers straight into the neighbour ranks' position buffers (peer memory over NVLink)
 * and publishes a per-step flag to every rank; the next step starts when all flags have arrived
 * (no host round trip, no NCCL call per step).  A skin violation on any rank halts all ranks at
 * the same step; the host then rebuilds the lists: slab-locally (ljmd_ctx_slab_halo / _rebuild_local below: two cell
 * layers exchanged with each neighbour, one all-reduce of the cell histogram) or globally (one NCCL all-gather of
 * the packed slabs, ljmd_ctx_slab_pack / _unpack, then ljmd_ctx_slab_rebuild re-sorts everything on every rank).
 *   export  : 128 bytes = two cudaIpcMemHandle_t (position buffers, flag block) and/or the raw
 *             device pointers (for contexts living in the same process)
 *   attach  : all_handles = world x 128 bytes in rank order (or all_raw_ptrs = world x 2) */
#define LJMD_MAX_RANKS 8
int ljmd_ctx_slab_export(ljmd_ctx *ctx, void *handles128, uint64_t *raw_ptrs2);
int ljmd_ctx_slab_attach(ljmd_ctx *ctx, int32_t rank, int32_t world, const void *all_handles,
                         const uint64_t *all_raw_ptrs);
/* First slot of every rank's slab in the current sorted order; bounds[world] = n. */
int ljmd_ctx_slab_bounds(ljmd_ctx *ctx, int32_t *bounds);
/* Owned range of every state array and the original atom indices -> dense device buffer [narr][cmax]
 * (narr = state arrays + 1; dev_buf NULL: query narr). */
int ljmd_ctx_slab_pack(ljmd_ctx *ctx, double *dev_buf, int32_t cmax, int32_t *narr);
/* All ranks' packed ranges [world][narr][cmax] -> state arrays: the complete state on this rank. */
int ljmd_ctx_slab_unpack(ljmd_ctx *ctx, const double *dev_buf, int32_t cmax);
/* r(t+dt) of all atoms from the complete state (start of a sharded run). */
int ljmd_ctx_slab_drift(ljmd_ctx *ctx);
/* Global re-sort + list rebuild for the next step's positions (needs the complete state) if a
 * halt / the skin check asked for it or force != 0; clears the halt. */
int ljmd_ctx_slab_rebuild(ljmd_ctx *ctx, int32_t force);
/* Up to max_steps steps (<= 32 of them are replayed as one CUDA graph); *halted != 0 => rebuild the lists (locally or
 * globally, see above), then go on.  A neighbour-list overflow or a peer whose flag never arrives (bounded wait) is
 * reported here as an error, with the state frozen at the last completed step.
 * obs4 = this rank's share of {epot, ekin, d_epot, dd_epot} of the last step, no tail corrections. */
int ljmd_ctx_slab_run(ljmd_ctx *ctx, int32_t max_steps, int32_t *steps_done, int32_t *halted,
                      double *obs4);

/* Slab-LOCAL rebuild (needs >= 2 ranks and >= 2 cell layers per rank): instead of the all-gather + global re-sort,
 * every rank exchanges two cell layers with each neighbour and re-sorts its own slab and halo layers only; the cell
 * table stays global through ONE all-reduce of the cell histogram.  Same slots, lists and results as the global
 * path, bit for bit.  Sequence (multi_gpu.SlabRunner.rebuild_local):
 *   halo(buf, cap, side, 0) x 2 -> [host: send / recv with the neighbours] -> halo(buf, cap, side, 1) x 2
 *   -> rebuild_local(0) -> [host: all-reduce(sum, int32) of the ncells ints at *cell_count_ptr] -> rebuild_local(1).
 * side 0 = towards / from the left neighbour, 1 = right.  Buffers: (halo_narr x cap) doubles, cap >= *halo_cap. */
int ljmd_ctx_slab_local_info(ljmd_ctx *ctx, uint64_t *cell_count_ptr, int32_t *ncells, int32_t *halo_cap,
                             int32_t *halo_narr);
int ljmd_ctx_slab_halo(ljmd_ctx *ctx, double *dev_buf, int32_t cap, int32_t side, int32_t unpack);
int ljmd_ctx_slab_rebuild_local(ljmd_ctx *ctx, int32_t phase);
/* Captures the CUDA graph that ljmd_ctx_slab_run replays for a chunk of max_steps (2 .. 32) steps now, so that the
 * first run of that length does not pay for the capture.  Optional. */
int ljmd_ctx_slab_prepare(ljmd_ctx *ctx, int32_t max_steps);

/* Self-check of the neighbour-list path: total momentum (sum of v) of the atoms handled by the last
 * step launch -- an invariant of the dynamics (a sharded run adds the ranks' shares). */
int ljmd_ctx_nlist_momentum(ljmd_ctx *ctx, double *p3);
/* A rank that cannot go on (an error from ljmd_ctx_slab_run / _rebuild) tells its peers: its per-step flag
 * then reads "done, skin violated", so their step kernels halt instead of waiting for it.  (The wait is
 * bounded anyway: LJMD_SLAB_SPIN_TIMEOUT_S seconds, default 10, then LJMD_ERR_CUDA.) */
int ljmd_ctx_slab_poison(ljmd_ctx *ctx);
/* Introspection used by tests and benchmarks. */
int ljmd_ctx_method(const ljmd_ctx *ctx);          /* resolved LJMD_METHOD_* */
int ljmd_ctx_cells_per_edge(const ljmd_ctx *ctx);  /* 0 for all-pairs */
/* Kernel launches issued by this context since creation (for bench.py's gpu_launches). */
int64_t ljmd_ctx_launch_count(const ljmd_ctx *ctx);
/* Device time (CUDA events on the context's stream, first kernel to last kernel) of the most
 * recent ljmd_ctx_step call, in milliseconds. */
int ljmd_ctx_last_step_ms(const ljmd_ctx *ctx, double *ms);
/* Device time of the force kernel(s) accumulated since the last reset, in milliseconds, and the
 * number of force evaluations it covers.  Timing is only recorded when enabled. */
int ljmd_ctx_force_timing(ljmd_ctx *ctx, int32_t enable, double *ms_total, int64_t *n_evals);

/* Cell-list integer state of the last force evaluation (LJMD_METHOD_CELLS only):
 * cell_of_atom[n] (cell id of each atom, original numbering), cell_start[ncell^3+1],
 * order[n] (original index of the k-th atom in cell-sorted order).  Any pointer may be NULL. */
int ljmd_ctx_download_cells(ljmd_ctx *ctx, int32_t *cell_of_atom, int32_t *cell_start,
                            int32_t *order);
/* Neighbour-list introspection (LJMD_METHOD_NLIST only).  Any pointer may be NULL.
 * skin: list radius is rc + skin; capacity: list slots per atom; rebuilds: list builds so far;
 * entries: total (ordered) list entries of the current list. */
int ljmd_ctx_nlist_info(ljmd_ctx *ctx, double *skin, int32_t *cells_per_edge, int32_t *capacity,
                        int64_t *rebuilds, int64_t *entries);
/* Neighbour list of the last build in ORIGINAL atom numbering: counts[n]; the neighbours of atom i
 * (ascending) at nbr[i*cap .. i*cap + counts[i]).  j is listed for i iff the minimum-image r_ij^2,
 * computed in the reference's operation order (lj_potential_energy.f90:119-129), is < (rc+skin)^2. */
int ljmd_ctx_download_nlist(ljmd_ctx *ctx, int32_t *counts, int32_t *nbr, int32_t cap);
/* Unordered candidate pairs distance-tested / in-cutoff pairs of the last force evaluation. */
int ljmd_ctx_pair_counts(ljmd_ctx *ctx, int64_t *n_candidates, int64_t *n_in_cutoff);

/* Measured FP64 FMA throughput of the device the context lives on (register-resident DFMA
 * chains), in TFLOP/s; the roofline denominator for the all-pairs kernel. */
int ljmd_measure_fp64_peak(int32_t device, double *tflops);
/* The same chains launched back to back for `seconds` (0.05 .. 10) and timed over the second half: the
 * throughput the FP64 pipe sustains once the clocks have settled under that load -- the denominator for
 * a kernel timed inside a long run (ljmd_measure_fp64_peak is a burst right after idle). */
int ljmd_measure_fp64_peak_sustained(int32_t device, double seconds, double *tflops);

/* ==========================================================================================
 * Trajectory consumers: the heavy loops of the reference's analysis script
 * scripts/md_one_run_analysis.py, on the arrays its read_rva (:345-397) returns -- row-major
 * [n_snap][n] per component, HOST pointers (copied to the device inside the call).
 * `device` < 0 = current device.  `kernel_ms` (may be NULL) receives the CUDA-event time of the kernel.
 * ========================================================================================== */

/* compute_rdf (:537-598), the pair loop :566-584: for every snapshot and every pair i < j,
 * d = r_j - r_i, d -= L*rint(d/L), r = sqrt(dx*dx + dy*dy + dz*dz); pairs with r < rmax are counted in
 * bin (int)(r/dr), dr = rmax/nbins.  hist[nbins] receives UNORDERED pair counts (the reference adds
 * 2.0 per pair, :584) and is bit-exact.  The snapshot / particle subsampling (:555-561) and the
 * normalisation (:586-597) are host work on <= nbins numbers (analysis.py).  Returns LJMD_ERR_ARG if a
 * pair with r < rmax lands in bin index nbins (the reference raises IndexError there). */
int ljmd_rdf_histogram(const double *rx, const double *ry, const double *rz, int32_t n_snap, int32_t n,
                       double box_length, int32_t nbins, double rmax, int64_t *hist, int32_t device,
                       double *kernel_ms);

/* compute_msd_tau_timeorig (:404-441, kind = LJMD_CORR_MSD, arrays = unwrapped positions) and
 * compute_vacf_tau_timeorig (:444-489, kind = LJMD_CORR_VACF, arrays = velocities): time-origin
 * averaged over t0 = 0, stride, 2*stride, ... < n_snap-1 and over particles.  max_lag < 0 or
 * > n_snap-1 means n_snap-1 (:420-422); origin_stride < 1 means 1 (:423).  out has
 * min(max_lag, n_snap-1)+1 entries (1 entry when n_snap < 2, :416-417 / :456-458). */
enum { LJMD_CORR_MSD = 0, LJMD_CORR_VACF = 1 };
int ljmd_time_correlation(const double *ax, const double *ay, const double *az, int32_t n_snap, int32_t n,
                          int32_t max_lag, int32_t origin_stride, int32_t kind, double *out, int32_t device,
                          double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* LJMD_H */
