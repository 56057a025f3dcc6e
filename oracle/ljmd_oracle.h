/*
 * ljmd_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT PATH).
 *
 * Plain-C restatement of the reference's MD hot path
 * (Ledicia/Molecular-Dynamics-Simulation---Lennard-Jones-monoatomic-fluid, Fortran 90).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the shipped CUDA library never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs
 *  and no Fortran compiler exists in this image (F4), so this oracle
 * cannot be checked against the reference's own build.  It is pinned only against
 * (i) analytic identities, (ii) the exactly-reproducible integer RNG stream and
 * (iii) the SURVEY.md Appendix-B anchors (a numpy emulation, not Fortran output).
 *
 * All citations are file:line relative to /root/reference/scripts/.
 */
#ifndef LJMD_ORACLE_H
#define LJMD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of type(sim_params), base/md_types.f90:27-50, as gfortran lays it out
 * (2 x int32 then 11 x real64; 96 bytes). */
typedef struct orc_params {
    int32_t n;
    int32_t num_cells;
    double box_length;
    double inv_box_length;
    double volume;
    double density;
    double dt;
    double dt_half;
    double dt_square_half;
    double rc;
    double rc_square;
    double length_scale;
    double inv_length_scale;
} orc_params;

/* Error codes: the reference `stop`s; the oracle returns the code instead. */
enum {
    ORC_OK = 0,
    ORC_ERR_N = 1,
    ORC_ERR_BOX = 2,
    ORC_ERR_VOLUME = 3,
    ORC_ERR_RC = 4,
    ORC_ERR_RC_SQUARE = 5,
    ORC_ERR_NULL = 6,
    ORC_ERR_RC_HALF_BOX = 7,
    ORC_ERR_DT = 8,
    ORC_ERR_EKIN = 9
};

/* base/md_types.f90:105-169 (init_params + compute_derived_params). */
int orc_params_init(orc_params *p, int32_t n, double box_length, double dt, double rc,
                    int32_t num_cells);

/* physics/geometry_pbc.f90:80-88 */
double orc_minimum_image(double dx, double box_length, double inv_box_length);
/* physics/geometry_pbc.f90:39-59 */
int orc_wrap_positions(int32_t n, double *rx, double *ry, double *rz, double box_length);

/* physics/lj_potential_energy.f90:46-225 */
int orc_compute_lj(const orc_params *p, const double *rx, const double *ry, const double *rz,
                   double *ax, double *ay, double *az, double *epot, double *d_epot,
                   double *dd_epot);
/* Same pair arithmetic; also returns the number of in-cutoff unordered pairs. */
int64_t orc_count_pairs_in_cutoff(const orc_params *p, const double *rx, const double *ry,
                                  const double *rz);
/* Tail constants only, physics/lj_potential_energy.f90:205-223. */
void orc_tail_corrections(const orc_params *p, double *tail_epot, double *tail_d_epot,
                          double *tail_dd_epot);

/* physics/verlet.f90:41-97 */
int orc_verlet_step(const orc_params *p, double *rx, double *ry, double *rz, double *vx,
                    double *vy, double *vz, double *ax, double *ay, double *az, double *epot,
                    double *ekin, double *d_epot, double *dd_epot);

/* Driver-side kinetic energy (single fused sum), md_simulation_program.f90:238-240 and
 * md_initial_config_program.f90:92. */
double orc_ekin_driver(int32_t n, const double *vx, const double *vy, const double *vz);

/* base/random_numbers.f90:48-116.  The Fortran SAVEd state is made explicit. */
typedef struct orc_rng {
    double ma[56]; /* 1-based, ma[1..55] */
    int32_t inext, inextp, iff;
} orc_rng;
void orc_rng_reset(orc_rng *g);
double orc_random_uniform(orc_rng *g, int32_t *seed);

/* md_initial_config_program.f90:132-187 */
int orc_build_fcc_lattice(const orc_params *p, double *rx, double *ry, double *rz);
/* md_initial_config_program.f90:196-209 */
void orc_assign_random_velocities(const orc_params *p, double *vx, double *vy, double *vz,
                                  orc_rng *g, int32_t *seed);
/* md_initial_config_program.f90:218-234 */
void orc_remove_com_velocity(const orc_params *p, double *vx, double *vy, double *vz);
/* md_initial_config_program.f90:244-266 */
int orc_rescale_velocities(const orc_params *p, double *vx, double *vy, double *vz,
                           double target_energy, double epot);
/* md_initial_config_program.f90:73-105: lattice + velocities + COM + force + rescale + force.
 * Leaves a(t=0) in ax..az; does NOT run the warm-up loop (:113-116). */
int orc_initial_config(const orc_params *p, int32_t seed, double target_total_energy,
                       double *rx, double *ry, double *rz, double *vx, double *vy, double *vz,
                       double *ax, double *ay, double *az, double *epot, double *ekin);

/* As above; use_cells evaluates forces with the cell-list oracle, temp_target > 0 replaces the
 * energy target by U(t=0) + 1.5*N*temp_target. */
int orc_initial_config_ex(const orc_params *p, int32_t seed, double target_total_energy,
                          double temp_target, int use_cells, double *rx, double *ry, double *rz,
                          double *vx, double *vy, double *vz, double *ax, double *ay, double *az,
                          double *epot, double *ekin);

/* md_simulation_program.f90:339-353: ru += MIC(r_new - r_prev). */
void orc_unwrap_update(const orc_params *p, const double *rx, const double *ry,
                       const double *rz, const double *rx_prev, const double *ry_prev,
                       const double *rz_prev, double *rux, double *ruy, double *ruz);

/* nsteps x (rx_prev copy, verlet_step, unwrap update) as in md_simulation_program.f90:300-356.
 * rux..ruz may be NULL (then only verlet_step is iterated, as in
 * md_initial_config_program.f90:113-116).  obs = {epot, ekin, d_epot, dd_epot} of the last step.
 * If series != NULL it receives 4 doubles per step. */
int orc_run_steps(const orc_params *p, int32_t nsteps, double *rx, double *ry, double *rz,
                  double *vx, double *vy, double *vz, double *ax, double *ay, double *az,
                  double *rux, double *ruy, double *ruz, double *obs, double *series);

/* ---------------------------------------------------------------------------------------
 * Cell-list oracle.  The reference has NO cell list; this defines the integer
 * cell assignment the CUDA counting sort must reproduce bit-exactly, and a force routine that
 * uses the same per-pair arithmetic as orc_compute_lj over the 27-cell stencil.
 *   ncell = floor(L / rc) per edge (>= 3 required), cell edge = L / ncell,
 *   c_x = min(ncell-1, max(0, (int)floor(x * (ncell * invL))))  (wrapped x may equal L exactly,
 *   see geometry_pbc.f90:53-57 note in SURVEY.md a3), cell id = (c_x * ncell + c_y) * ncell + c_z,
 *   atoms inside a cell keep ascending original index (stable counting sort).
 * ------------------------------------------------------------------------------------- */
int32_t orc_cells_per_edge(const orc_params *p);
int orc_cell_assign(const orc_params *p, const double *rx, const double *ry, const double *rz,
                    int32_t *cell_of_atom);
/* cell_start has ncell^3+1 entries; order[k] = original index of the k-th atom in sorted order. */
int orc_cell_sort(const orc_params *p, const int32_t *cell_of_atom, int32_t *cell_start,
                  int32_t *order);
/* The same definitions for an arbitrary interaction range (the neighbour-list path bins with
 * range = rc + skin); orc_cells_per_edge(p) == orc_cells_per_edge_range(p, p->rc). */
int32_t orc_cells_per_edge_range(const orc_params *p, double range);
int orc_cell_assign_range(const orc_params *p, double range, const double *rx, const double *ry,
                          const double *rz, int32_t *cell_of_atom);
int orc_cell_sort_range(const orc_params *p, double range, const int32_t *cell_of_atom,
                        int32_t *cell_start, int32_t *order);
/* Verlet neighbour list: j is listed for i iff the minimum-image r_ij^2 (operation order of
 * lj_potential_energy.f90:119-129) is < range^2.  counts[n]; neighbours of i ascending at
 * nbr[i*cap ...].  Returns ORC_ERR_N if an atom has more than cap neighbours. */
int orc_neighbor_list(const orc_params *p, double range, const double *rx, const double *ry,
                      const double *rz, int32_t cap, int32_t *counts, int32_t *nbr);
int orc_compute_lj_cells(const orc_params *p, const double *rx, const double *ry,
                         const double *rz, double *ax, double *ay, double *az, double *epot,
                         double *d_epot, double *dd_epot, int64_t *n_candidates,
                         int64_t *n_in_cutoff);
int orc_verlet_step_cells(const orc_params *p, double *rx, double *ry, double *rz, double *vx,
                          double *vy, double *vz, double *ax, double *ay, double *az,
                          double *epot, double *ekin, double *d_epot, double *dd_epot);

/* ---------------------------------------------------------------------------------------
 * Consumers of the per-step reductions: stats/md_means.f90:192-366, stats/stats_math.f90:56-61,
 * physics/thermodynamic_coefs.f90:87-205.  Quantity order of sum / sum_sq / mean / std:
 * U, K, E, T, P, 1/K, dU, ddU, dU/K, dU^2/K, ddU/K.
 * ------------------------------------------------------------------------------------- */
typedef struct orc_means {
    int32_t n, n_samples;
    double volume;
    double sum[11], sum_sq[11];
} orc_means;
void orc_means_init(orc_means *m, int32_t n, double volume);
int orc_means_add(orc_means *m, double epot, double ekin, double virial, double d_epot,
                  double dd_epot, double *temp_inst, double *press_inst);
int orc_means_get(const orc_means *m, double *mean11, double *std11);
/* out[16] = temperature, pressure, Ca_v, Ce_v, gamma, K_S_aux, K_S, K_S_inv, K_T, K_T_inv, Ca_p, Ce_p,
 * alpha_E1, alpha_E2, alpha_S, alpha_P */
int orc_thermo_coefs(const orc_params *p, double ekin_mean, double press_mean,
                     double ekinInv_mean, double d_epot_mean, double dd_epot_mean,
                     double d_epot_ekinInv_mean, double d_epot2_ekinInv_mean, double *out16);

#ifdef __cplusplus
}
#endif
#endif
