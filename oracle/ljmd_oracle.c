/*
 * ljmd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT PATH).
 *
 * Line-by-line plain-C restatement of the reference Fortran hot path.  See ljmd_oracle.h for
 * the "parity unpinned" statement.  Build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math
 * (the shipped reference flags are gfortran -O2 on baseline x86-64: no FMA contraction, no
 * reassociation, in-order sum(); build/one_run/compile_md_simulation.sh:19-31).
 *
 * Citations are file:line relative to /root/reference/scripts/.
 */
#include "ljmd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* base/md_types.f90:22 */
static const double ORC_PI = 3.1415926535897932384626433832795;

/* ------------------------------------------------------------------------------------------
 * base/md_types.f90:105-169
 * ---------------------------------------------------------------------------------------- */
int orc_params_init(orc_params *p, int32_t n, double box_length, double dt, double rc,
                    int32_t num_cells)
{
    if (!p) return ORC_ERR_NULL;
    memset(p, 0, sizeof(*p));
    p->n = n;                   /* :113 */
    p->box_length = box_length; /* :114 */
    p->dt = dt;                 /* :115 */
    p->rc = rc;                 /* :116 */
    p->num_cells = num_cells;   /* :117 */

    /* compute_derived_params, :132-169 */
    if (p->box_length > 0.0) {
        p->inv_box_length = 1.0 / p->box_length;                      /* :137 */
        p->volume = (p->box_length * p->box_length) * p->box_length;  /* :138, L**3 */
        if (p->n > 0) p->density = (double)p->n / p->volume;          /* :141 */
    } else {
        return ORC_ERR_BOX; /* :143 */
    }
    if (p->rc > 0.0) {
        p->rc_square = p->rc * p->rc; /* :148 */
    } else {
        return ORC_ERR_RC; /* :150 */
    }
    if (p->rc >= 0.5 * p->box_length) return ORC_ERR_RC_HALF_BOX; /* :152 */
    if (p->dt > 0.0) {
        p->dt_half = 0.5 * p->dt;               /* :158 */
        p->dt_square_half = p->dt_half * p->dt; /* :159 */
    } else {
        return ORC_ERR_DT; /* :161 */
    }
    if (p->length_scale > 0.0) p->inv_length_scale = 1.0 / p->length_scale; /* :165-167 */
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * physics/geometry_pbc.f90:80-88.  dnint = round half away from zero = C round().
 * ---------------------------------------------------------------------------------------- */
double orc_minimum_image(double dx, double box_length, double inv_box_length)
{
    return dx - box_length * round(dx * inv_box_length); /* :86 */
}

/* physics/geometry_pbc.f90:39-59.  Fortran floor() returns a default INTEGER that is converted
 * back to real64; identical to C floor() for |x/L| < 2^31. */
int orc_wrap_positions(int32_t n, double *rx, double *ry, double *rz, double box_length)
{
    if (box_length <= 0.0) return ORC_ERR_BOX; /* :47 */
    const double invL = 1.0 / box_length;      /* :51 */
    for (int32_t i = 0; i < n; ++i) {          /* :53-57 */
        rx[i] = rx[i] - box_length * floor(rx[i] * invL);
        ry[i] = ry[i] - box_length * floor(ry[i] * invL);
        rz[i] = rz[i] - box_length * floor(rz[i] * invL);
    }
    return ORC_OK;
}

/* physics/lj_potential_energy.f90:205-223.  Integer powers expand to multiplications:
 * rc**3 = (rc*rc)*rc, rc**6 = (rc**3)*(rc**3). */
void orc_tail_corrections(const orc_params *p, double *tail_epot, double *tail_d_epot,
                          double *tail_dd_epot)
{
    const double npd = (double)p->n;                                   /* :206 */
    const double rc3 = (p->rc * p->rc) * p->rc;
    const double rc6 = rc3 * rc3;
    const double tail_factor = 8.0 * ORC_PI * (npd * npd) / (p->volume * rc3); /* :208 */
    *tail_epot = tail_factor * ((1.0 / (3.0 * rc6)) - 1.0) / 3.0;              /* :210 */
    *tail_d_epot = 2.0 * tail_factor * (-2.0 / (3.0 * rc6) + 1.0);             /* :211 */
    *tail_dd_epot = 2.0 * tail_factor * (26.0 / (3.0 * rc6) - 7.0);            /* :212 */
}

static int orc_check_lj_params(const orc_params *p)
{
    if (p->n <= 0) return ORC_ERR_N;                    /* :77 */
    if (p->box_length <= 0.0) return ORC_ERR_BOX;       /* :78 */
    if (p->volume <= 0.0) return ORC_ERR_VOLUME;        /* :79 */
    if (p->rc <= 0.0) return ORC_ERR_RC;                /* :80 */
    if (p->rc_square <= 0.0) return ORC_ERR_RC_SQUARE;  /* :81 */
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * physics/lj_potential_energy.f90:46-225
 * ---------------------------------------------------------------------------------------- */
int orc_compute_lj(const orc_params *p, const double *rx, const double *ry, const double *rz,
                   double *ax, double *ay, double *az, double *epot_out, double *d_epot_out,
                   double *dd_epot_out)
{
    if (!p) return ORC_ERR_NULL;
    int rc_ = orc_check_lj_params(p);
    if (rc_) return rc_;
    if (!rx || !ry || !rz || !ax || !ay || !az) return ORC_ERR_NULL; /* :82 */

    const int32_t n = p->n;
    const double L = p->box_length, invL = p->inv_box_length, rc2 = p->rc_square;
    double epot = 0.0, d_epot = 0.0, dd_epot = 0.0; /* :87-89 */
    for (int32_t i = 0; i < n; ++i) { ax[i] = 0.0; ay[i] = 0.0; az[i] = 0.0; } /* :91-93 */

    for (int32_t i = 0; i < n - 1; ++i) { /* :109 */
        const double xi = rx[i], yi = ry[i], zi = rz[i]; /* :112-114 */
        for (int32_t j = i + 1; j < n; ++j) {            /* :116 */
            double dx = xi - rx[j];                      /* :119-121 */
            double dy = yi - ry[j];
            double dz = zi - rz[j];
            dx = orc_minimum_image(dx, L, invL);         /* :124-126 */
            dy = orc_minimum_image(dy, L, invL);
            dz = orc_minimum_image(dz, L, invL);
            const double rij2 = dx * dx + dy * dy + dz * dz; /* :129 */
            if (rij2 < rc2) {                                /* :132 strict */
                const double inv_r2 = 1.0 / rij2;                 /* :135 */
                const double inv_r6 = inv_r2 * inv_r2 * inv_r2;   /* :136 */
                const double inv_r12 = inv_r6 * inv_r6;           /* :137 */
                epot = epot + (inv_r12 - inv_r6);                 /* :140 */
                const double dU_r = (-2.0 * inv_r12 + inv_r6);    /* :143 */
                const double fx = -dU_r * dx * inv_r2;            /* :148-150 */
                const double fy = -dU_r * dy * inv_r2;
                const double fz = -dU_r * dz * inv_r2;
                ax[i] = ax[i] + fx; ay[i] = ay[i] + fy; az[i] = az[i] + fz; /* :153-155 */
                ax[j] = ax[j] - fx; ay[j] = ay[j] - fy; az[j] = az[j] - fz; /* :157-159 */
                d_epot = d_epot + dU_r;                                     /* :177 */
                dd_epot = dd_epot + (26.0 * inv_r12 - 7.0 * inv_r6);        /* :178 */
            }
        }
    }

    epot = 4.0 * epot; /* :188 */
    for (int32_t i = 0; i < n; ++i) { ax[i] = 24.0 * ax[i]; } /* :189-191 */
    for (int32_t i = 0; i < n; ++i) { ay[i] = 24.0 * ay[i]; }
    for (int32_t i = 0; i < n; ++i) { az[i] = 24.0 * az[i]; }
    d_epot = 24.0 * d_epot;   /* :192 */
    dd_epot = 24.0 * dd_epot; /* :193 */

    double te, td, tdd; /* use_tail_corrections = .true., :36 */
    orc_tail_corrections(p, &te, &td, &tdd);
    *epot_out = epot + te;       /* :221 */
    *d_epot_out = d_epot + td;   /* :222 */
    *dd_epot_out = dd_epot + tdd; /* :223 */
    return ORC_OK;
}

int64_t orc_count_pairs_in_cutoff(const orc_params *p, const double *rx, const double *ry,
                                  const double *rz)
{
    const int32_t n = p->n;
    const double L = p->box_length, invL = p->inv_box_length, rc2 = p->rc_square;
    int64_t cnt = 0;
    for (int32_t i = 0; i < n - 1; ++i) {
        for (int32_t j = i + 1; j < n; ++j) {
            const double dx = orc_minimum_image(rx[i] - rx[j], L, invL);
            const double dy = orc_minimum_image(ry[i] - ry[j], L, invL);
            const double dz = orc_minimum_image(rz[i] - rz[j], L, invL);
            const double rij2 = dx * dx + dy * dy + dz * dz;
            if (rij2 < rc2) ++cnt;
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * physics/verlet.f90:41-97
 * ---------------------------------------------------------------------------------------- */
static void orc_drift_wrap_halfkick(const orc_params *p, double *rx, double *ry, double *rz,
                                    double *vx, double *vy, double *vz, const double *ax,
                                    const double *ay, const double *az)
{
    const int32_t n = p->n;
    const double dt = p->dt, dth = p->dt_half, dtsqh = p->dt_square_half;
    for (int32_t i = 0; i < n; ++i) rx[i] = rx[i] + vx[i] * dt + ax[i] * dtsqh; /* :58 */
    for (int32_t i = 0; i < n; ++i) ry[i] = ry[i] + vy[i] * dt + ay[i] * dtsqh; /* :59 */
    for (int32_t i = 0; i < n; ++i) rz[i] = rz[i] + vz[i] * dt + az[i] * dtsqh; /* :60 */
    orc_wrap_positions(n, rx, ry, rz, p->box_length);                           /* :63 */
    for (int32_t i = 0; i < n; ++i) vx[i] = vx[i] + ax[i] * dth; /* :72-74 */
    for (int32_t i = 0; i < n; ++i) vy[i] = vy[i] + ay[i] * dth;
    for (int32_t i = 0; i < n; ++i) vz[i] = vz[i] + az[i] * dth;
}

static double orc_halfkick_ekin(const orc_params *p, double *vx, double *vy, double *vz,
                                const double *ax, const double *ay, const double *az)
{
    const int32_t n = p->n;
    const double dth = p->dt_half;
    for (int32_t i = 0; i < n; ++i) vx[i] = vx[i] + ax[i] * dth; /* :86-88 */
    for (int32_t i = 0; i < n; ++i) vy[i] = vy[i] + ay[i] * dth;
    for (int32_t i = 0; i < n; ++i) vz[i] = vz[i] + az[i] * dth;
    double sx = 0.0, sy = 0.0, sz = 0.0; /* :93-95: three separate in-order sums */
    for (int32_t i = 0; i < n; ++i) sx = sx + vx[i] * vx[i];
    for (int32_t i = 0; i < n; ++i) sy = sy + vy[i] * vy[i];
    for (int32_t i = 0; i < n; ++i) sz = sz + vz[i] * vz[i];
    return 0.5 * (sx + sy + sz);
}

int orc_verlet_step(const orc_params *p, double *rx, double *ry, double *rz, double *vx,
                    double *vy, double *vz, double *ax, double *ay, double *az, double *epot,
                    double *ekin, double *d_epot, double *dd_epot)
{
    if (!p) return ORC_ERR_NULL;
    if (p->n <= 0) return ORC_ERR_N; /* :51 */
    if (!rx || !ry || !rz || !vx || !vy || !vz || !ax || !ay || !az) return ORC_ERR_NULL; /* :52 */
    orc_drift_wrap_halfkick(p, rx, ry, rz, vx, vy, vz, ax, ay, az);
    int rc_ = orc_compute_lj(p, rx, ry, rz, ax, ay, az, epot, d_epot, dd_epot); /* :80 */
    if (rc_) return rc_;
    *ekin = orc_halfkick_ekin(p, vx, vy, vz, ax, ay, az);
    return ORC_OK;
}

/* md_simulation_program.f90:238-240; md_initial_config_program.f90:92,105,254 */
double orc_ekin_driver(int32_t n, const double *vx, const double *vy, const double *vz)
{
    double s = 0.0;
    for (int32_t i = 0; i < n; ++i) s = s + (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]);
    return 0.5 * s;
}

/* ------------------------------------------------------------------------------------------
 * base/random_numbers.f90:48-116 (Knuth subtractive generator held in real64).
 * ---------------------------------------------------------------------------------------- */
void orc_rng_reset(orc_rng *g) { memset(g, 0, sizeof(*g)); }

double orc_random_uniform(orc_rng *g, int32_t *seed)
{
    const double mbig = 4.0e6, mseed = 1618033.0, mz = 0.0; /* :54-56 */
    const double fac = 1.0 / mbig;                          /* :57 */
    double mj, mk;
    if (*seed <= 0 || g->iff == 0) { /* :72 */
        g->iff = 1;
        mj = fabs(mseed - fabs((double)*seed)); /* :75 */
        mj = fmod(mj, mbig);                    /* :76 */
        g->ma[55] = mj;
        mk = 1.0;
        for (int32_t i = 1; i <= 54; ++i) { /* :80-86 */
            const int32_t ii = (21 * i) % 55;
            g->ma[ii] = mk;
            mk = mj - mk;
            if (mk < mz) mk = mk + mbig;
            mj = g->ma[ii];
        }
        for (int32_t k = 1; k <= 4; ++k) { /* :88-93 */
            for (int32_t i = 1; i <= 55; ++i) {
                g->ma[i] = g->ma[i] - g->ma[1 + (i + 30) % 55];
                if (g->ma[i] < mz) g->ma[i] = g->ma[i] + mbig;
            }
        }
        g->inext = 0;   /* :95 */
        g->inextp = 31; /* :96 */
        *seed = 1;      /* :97 */
    }
    g->inext = g->inext + 1; /* :104-108 */
    if (g->inext == 56) g->inext = 1;
    g->inextp = g->inextp + 1;
    if (g->inextp == 56) g->inextp = 1;
    mj = g->ma[g->inext] - g->ma[g->inextp]; /* :110 */
    if (mj < mz) mj = mj + mbig;
    g->ma[g->inext] = mj;
    return mj * fac; /* :114 */
}

/* ------------------------------------------------------------------------------------------
 * md_initial_config_program.f90:132-187: loops ix (outer), iy, iz (inner); 4 basis atoms.
 * ---------------------------------------------------------------------------------------- */
int orc_build_fcc_lattice(const orc_params *p, double *rx, double *ry, double *rz)
{
    const double a = p->box_length / (double)p->num_cells; /* :73 */
    int32_t idx = 0;
    for (int32_t ix = 0; ix < p->num_cells; ++ix)
        for (int32_t iy = 0; iy < p->num_cells; ++iy)
            for (int32_t iz = 0; iz < p->num_cells; ++iz) {
                const double x0 = (double)ix * a, y0 = (double)iy * a, z0 = (double)iz * a;
                rx[idx] = x0;           ry[idx] = y0;           rz[idx] = z0;           ++idx;
                rx[idx] = x0;           ry[idx] = y0 + 0.5 * a; rz[idx] = z0 + 0.5 * a; ++idx;
                rx[idx] = x0 + 0.5 * a; ry[idx] = y0;           rz[idx] = z0 + 0.5 * a; ++idx;
                rx[idx] = x0 + 0.5 * a; ry[idx] = y0 + 0.5 * a; rz[idx] = z0;           ++idx;
            }
    return idx == p->n ? ORC_OK : ORC_ERR_N; /* :185 */
}

/* md_initial_config_program.f90:196-209: draw order vx, vy, vz per particle. */
void orc_assign_random_velocities(const orc_params *p, double *vx, double *vy, double *vz,
                                  orc_rng *g, int32_t *seed)
{
    for (int32_t i = 0; i < p->n; ++i) {
        vx[i] = orc_random_uniform(g, seed) - 0.5;
        vy[i] = orc_random_uniform(g, seed) - 0.5;
        vz[i] = orc_random_uniform(g, seed) - 0.5;
    }
}

/* md_initial_config_program.f90:218-234 */
void orc_remove_com_velocity(const orc_params *p, double *vx, double *vy, double *vz)
{
    const int32_t n = p->n;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int32_t i = 0; i < n; ++i) sx = sx + vx[i];
    for (int32_t i = 0; i < n; ++i) sy = sy + vy[i];
    for (int32_t i = 0; i < n; ++i) sz = sz + vz[i];
    const double cx = sx / (double)n, cy = sy / (double)n, cz = sz / (double)n; /* :226-228 */
    for (int32_t i = 0; i < n; ++i) vx[i] = vx[i] - cx; /* :230-232 */
    for (int32_t i = 0; i < n; ++i) vy[i] = vy[i] - cy;
    for (int32_t i = 0; i < n; ++i) vz[i] = vz[i] - cz;
}

/* md_initial_config_program.f90:244-266 */
int orc_rescale_velocities(const orc_params *p, double *vx, double *vy, double *vz,
                           double target_energy, double epot)
{
    const int32_t n = p->n;
    const double ekin_old = orc_ekin_driver(n, vx, vy, vz); /* :254 */
    const double ekin_new = target_energy - epot;           /* :255 */
    if (ekin_new <= 0.0) return ORC_ERR_EKIN;               /* :257 */
    if (ekin_old <= 0.0) return ORC_ERR_EKIN;               /* :258 */
    const double scale = sqrt(ekin_new / ekin_old);         /* :260 */
    for (int32_t i = 0; i < n; ++i) vx[i] = vx[i] * scale;  /* :262-264 */
    for (int32_t i = 0; i < n; ++i) vy[i] = vy[i] * scale;
    for (int32_t i = 0; i < n; ++i) vz[i] = vz[i] * scale;
    return ORC_OK;
}

/* md_initial_config_program.f90:73-105.  use_cells != 0 evaluates the two force calls with the
 * cell-list oracle (same per-pair arithmetic; for N where O(N^2) is too slow).  If
 * temp_target > 0 the target total energy is replaced by U(t=0) + 1.5*N*temp_target (used for
 * the rho=0.8 benchmark configurations, which the reference does not ship). */
static int orc_force_any(const orc_params *p, int use_cells, const double *rx, const double *ry,
                         const double *rz, double *ax, double *ay, double *az, double *epot)
{
    double d1, d2;
    if (use_cells)
        return orc_compute_lj_cells(p, rx, ry, rz, ax, ay, az, epot, &d1, &d2, NULL, NULL);
    return orc_compute_lj(p, rx, ry, rz, ax, ay, az, epot, &d1, &d2);
}

int orc_initial_config_ex(const orc_params *p, int32_t seed, double target_total_energy,
                          double temp_target, int use_cells, double *rx, double *ry, double *rz,
                          double *vx, double *vy, double *vz, double *ax, double *ay, double *az,
                          double *epot, double *ekin)
{
    int rc_ = orc_build_fcc_lattice(p, rx, ry, rz); /* :74 */
    if (rc_) return rc_;
    orc_rng g;
    orc_rng_reset(&g);
    int32_t s = seed;                                     /* :82 (reference: -12345) */
    orc_assign_random_velocities(p, vx, vy, vz, &g, &s);  /* :83 */
    orc_remove_com_velocity(p, vx, vy, vz);               /* :84 */
    rc_ = orc_force_any(p, use_cells, rx, ry, rz, ax, ay, az, epot); /* :91 */
    if (rc_) return rc_;
    if (temp_target > 0.0) target_total_energy = *epot + 1.5 * (double)p->n * temp_target;
    rc_ = orc_rescale_velocities(p, vx, vy, vz, target_total_energy, *epot); /* :101 */
    if (rc_) return rc_;
    rc_ = orc_force_any(p, use_cells, rx, ry, rz, ax, ay, az, epot); /* :104 */
    if (rc_) return rc_;
    *ekin = orc_ekin_driver(p->n, vx, vy, vz); /* :105 */
    return ORC_OK;
}

int orc_initial_config(const orc_params *p, int32_t seed, double target_total_energy,
                       double *rx, double *ry, double *rz, double *vx, double *vy, double *vz,
                       double *ax, double *ay, double *az, double *epot, double *ekin)
{
    return orc_initial_config_ex(p, seed, target_total_energy, 0.0, 0, rx, ry, rz, vx, vy, vz,
                                 ax, ay, az, epot, ekin);
}

/* md_simulation_program.f90:339-353 */
void orc_unwrap_update(const orc_params *p, const double *rx, const double *ry,
                       const double *rz, const double *rx_prev, const double *ry_prev,
                       const double *rz_prev, double *rux, double *ruy, double *ruz)
{
    const double L = p->box_length, invL = p->inv_box_length;
    for (int32_t i = 0; i < p->n; ++i) {
        double dx = rx[i] - rx_prev[i]; /* :341-343 */
        double dy = ry[i] - ry_prev[i];
        double dz = rz[i] - rz_prev[i];
        dx = dx - L * round(dx * invL); /* :345-347 */
        dy = dy - L * round(dy * invL);
        dz = dz - L * round(dz * invL);
        rux[i] = rux[i] + dx; /* :349-351 */
        ruy[i] = ruy[i] + dy;
        ruz[i] = ruz[i] + dz;
    }
}

/* md_simulation_program.f90:300-356 (with ru) / md_initial_config_program.f90:113-116 (without) */
int orc_run_steps(const orc_params *p, int32_t nsteps, double *rx, double *ry, double *rz,
                  double *vx, double *vy, double *vz, double *ax, double *ay, double *az,
                  double *rux, double *ruy, double *ruz, double *obs, double *series)
{
    const int32_t n = p->n;
    double *prev = NULL;
    if (rux) {
        prev = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        if (!prev) return ORC_ERR_NULL;
    }
    double o[4] = {0, 0, 0, 0};
    for (int32_t s = 0; s < nsteps; ++s) {
        if (rux) { /* :303-305 */
            memcpy(prev, rx, sizeof(double) * n);
            memcpy(prev + n, ry, sizeof(double) * n);
            memcpy(prev + 2 * n, rz, sizeof(double) * n);
        }
        int rc_ = orc_verlet_step(p, rx, ry, rz, vx, vy, vz, ax, ay, az, &o[0], &o[1], &o[2],
                                  &o[3]); /* :322 */
        if (rc_) { free(prev); return rc_; }
        if (rux) orc_unwrap_update(p, rx, ry, rz, prev, prev + n, prev + 2 * n, rux, ruy, ruz);
        if (series) memcpy(series + 4 * (size_t)s, o, sizeof(o));
    }
    if (obs) memcpy(obs, o, sizeof(o));
    free(prev);
    return ORC_OK;
}

/* ==========================================================================================
 * Cell-list oracle (no reference counterpart; definition in ljmd_oracle.h).
 * ======================================================================================== */
int32_t orc_cells_per_edge_range(const orc_params *p, double range)
{
    /* Largest ncell with cell edge L/ncell >= range, backed off by one if the edge is within
     * 1e-12 relative of range so that a rounding in floor(x*scale) can never separate an in-range
     * pair by two cells. */
    int32_t nc = (int32_t)floor(p->box_length / range);
    if (nc > 0 && (double)nc * range * (1.0 + 1e-12) > p->box_length) nc -= 1;
    if (nc > 128) nc = 128; /* bounded cell table: at most 128^3 cells */
    return nc;
}

int32_t orc_cells_per_edge(const orc_params *p) { return orc_cells_per_edge_range(p, p->rc); }

static inline int32_t orc_cell_coord(double x, double L, double invL, double scale, int32_t nc)
{
    /* out-of-box coordinates (unwrapped input, geometry_pbc.f90:18-19) are wrapped for binning */
    if (x < 0.0 || x > L) x = x - L * floor(x * invL);
    int32_t c = (int32_t)floor(x * scale);
    if (c < 0) c = 0;
    if (c > nc - 1) c = nc - 1; /* wrapped x may equal L exactly */
    return c;
}

int orc_cell_assign_range(const orc_params *p, double range, const double *rx, const double *ry,
                          const double *rz, int32_t *cell_of_atom)
{
    const int32_t nc = orc_cells_per_edge_range(p, range);
    if (nc < 3) return ORC_ERR_RC;
    const double scale = (double)nc * p->inv_box_length;
    const double L = p->box_length, invL = p->inv_box_length;
    for (int32_t i = 0; i < p->n; ++i) {
        const int32_t cx = orc_cell_coord(rx[i], L, invL, scale, nc);
        const int32_t cy = orc_cell_coord(ry[i], L, invL, scale, nc);
        const int32_t cz = orc_cell_coord(rz[i], L, invL, scale, nc);
        cell_of_atom[i] = (cx * nc + cy) * nc + cz;
    }
    return ORC_OK;
}

int orc_cell_assign(const orc_params *p, const double *rx, const double *ry, const double *rz,
                    int32_t *cell_of_atom)
{
    return orc_cell_assign_range(p, p->rc, rx, ry, rz, cell_of_atom);
}

int orc_cell_sort_range(const orc_params *p, double range, const int32_t *cell_of_atom,
                        int32_t *cell_start, int32_t *order)
{
    const int32_t nc = orc_cells_per_edge_range(p, range);
    if (nc < 3) return ORC_ERR_RC;
    const int32_t ncells = nc * nc * nc;
    memset(cell_start, 0, sizeof(int32_t) * ((size_t)ncells + 1));
    for (int32_t i = 0; i < p->n; ++i) cell_start[cell_of_atom[i] + 1] += 1;
    for (int32_t c = 0; c < ncells; ++c) cell_start[c + 1] += cell_start[c];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncells);
    if (!fill) return ORC_ERR_NULL;
    memcpy(fill, cell_start, sizeof(int32_t) * (size_t)ncells);
    for (int32_t i = 0; i < p->n; ++i) order[fill[cell_of_atom[i]]++] = i; /* stable */
    free(fill);
    return ORC_OK;
}

int orc_cell_sort(const orc_params *p, const int32_t *cell_of_atom, int32_t *cell_start,
                  int32_t *order)
{
    return orc_cell_sort_range(p, p->rc, cell_of_atom, cell_start, order);
}

static int orc_cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* Verlet neighbour list (no reference counterpart): j is a neighbour of i iff the minimum-image
 * r_ij^2, computed exactly as lj_potential_energy.f90:119-129 computes it, is < range^2.
 * Full list (both i->j and j->i), each atom's neighbours ascending.  counts[n], nbr[n*cap]. */
int orc_neighbor_list(const orc_params *p, double range, const double *rx, const double *ry,
                      const double *rz, int32_t cap, int32_t *counts, int32_t *nbr)
{
    const int32_t n = p->n;
    const int32_t nc = orc_cells_per_edge_range(p, range);
    if (nc < 3) return ORC_ERR_RC;
    const int32_t ncells = nc * nc * nc;
    int32_t *cell = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *start = (int32_t *)malloc(sizeof(int32_t) * ((size_t)ncells + 1));
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!cell || !start || !order) { free(cell); free(start); free(order); return ORC_ERR_NULL; }
    orc_cell_assign_range(p, range, rx, ry, rz, cell);
    orc_cell_sort_range(p, range, cell, start, order);
    const double L = p->box_length, invL = p->inv_box_length, range2 = range * range;
    int rc_ = ORC_OK;
    for (int32_t i = 0; i < n && rc_ == ORC_OK; ++i) {
        const int32_t ci = cell[i];
        const int32_t cx = ci / (nc * nc), cy = (ci / nc) % nc, cz = ci % nc;
        int32_t cnt = 0;
        /* visit every distinct neighbour cell once (nc == 3: the 27 offsets cover the box once) */
        for (int32_t ox = -1; ox <= 1; ++ox)
            for (int32_t oy = -1; oy <= 1; ++oy)
                for (int32_t oz = -1; oz <= 1; ++oz) {
                    const int32_t nx = (cx + ox + nc) % nc, ny = (cy + oy + nc) % nc,
                                  nz = (cz + oz + nc) % nc;
                    const int32_t cj = (nx * nc + ny) * nc + nz;
                    for (int32_t k = start[cj]; k < start[cj + 1]; ++k) {
                        const int32_t j = order[k];
                        if (j == i) continue;
                        double dx = rx[i] - rx[j], dy = ry[i] - ry[j], dz = rz[i] - rz[j];
                        dx = orc_minimum_image(dx, L, invL);
                        dy = orc_minimum_image(dy, L, invL);
                        dz = orc_minimum_image(dz, L, invL);
                        const double rij2 = dx * dx + dy * dy + dz * dz;
                        if (rij2 < range2) {
                            if (cnt >= cap) { rc_ = ORC_ERR_N; break; }
                            nbr[(size_t)i * cap + cnt++] = j;
                        }
                    }
                }
        counts[i] = cnt;
        qsort(nbr + (size_t)i * cap, (size_t)cnt, sizeof(int32_t), orc_cmp_i32);
    }
    free(cell); free(start); free(order);
    return rc_;
}

int orc_compute_lj_cells(const orc_params *p, const double *rx, const double *ry,
                         const double *rz, double *ax, double *ay, double *az,
                         double *epot_out, double *d_epot_out, double *dd_epot_out,
                         int64_t *n_candidates, int64_t *n_in_cutoff)
{
    if (!p) return ORC_ERR_NULL;
    int rc_ = orc_check_lj_params(p);
    if (rc_) return rc_;
    const int32_t n = p->n;
    const int32_t nc = orc_cells_per_edge(p);
    if (nc < 3) return ORC_ERR_RC;
    const int32_t ncells = nc * nc * nc;
    int32_t *cell = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *start = (int32_t *)malloc(sizeof(int32_t) * ((size_t)ncells + 1));
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!cell || !start || !order) { free(cell); free(start); free(order); return ORC_ERR_NULL; }
    orc_cell_assign(p, rx, ry, rz, cell);
    orc_cell_sort(p, cell, start, order);

    const double L = p->box_length, invL = p->inv_box_length, rc2 = p->rc_square;
    double epot = 0.0, d_epot = 0.0, dd_epot = 0.0;
    int64_t ncand = 0, ncut = 0;
    for (int32_t i = 0; i < n; ++i) { ax[i] = 0.0; ay[i] = 0.0; az[i] = 0.0; }

    /* Each unordered pair is visited once, from the atom with the smaller original index,
     * with exactly the per-pair arithmetic of orc_compute_lj (d = r_i - r_j, i < j). */
    for (int32_t i = 0; i < n; ++i) {
        const int32_t ci = cell[i];
        const int32_t cx = ci / (nc * nc), cy = (ci / nc) % nc, cz = ci % nc;
        const double xi = rx[i], yi = ry[i], zi = rz[i];
        for (int32_t ox = -1; ox <= 1; ++ox)
            for (int32_t oy = -1; oy <= 1; ++oy)
                for (int32_t oz = -1; oz <= 1; ++oz) {
                    const int32_t nx = (cx + ox + nc) % nc, ny = (cy + oy + nc) % nc,
                                  nz = (cz + oz + nc) % nc;
                    const int32_t cj = (nx * nc + ny) * nc + nz;
                    for (int32_t k = start[cj]; k < start[cj + 1]; ++k) {
                        const int32_t j = order[k];
                        if (j <= i) continue;
                        ++ncand;
                        double dx = xi - rx[j], dy = yi - ry[j], dz = zi - rz[j];
                        dx = orc_minimum_image(dx, L, invL);
                        dy = orc_minimum_image(dy, L, invL);
                        dz = orc_minimum_image(dz, L, invL);
                        const double rij2 = dx * dx + dy * dy + dz * dz;
                        if (rij2 < rc2) {
                            ++ncut;
                            const double inv_r2 = 1.0 / rij2;
                            const double inv_r6 = inv_r2 * inv_r2 * inv_r2;
                            const double inv_r12 = inv_r6 * inv_r6;
                            epot = epot + (inv_r12 - inv_r6);
                            const double dU_r = (-2.0 * inv_r12 + inv_r6);
                            const double fx = -dU_r * dx * inv_r2;
                            const double fy = -dU_r * dy * inv_r2;
                            const double fz = -dU_r * dz * inv_r2;
                            ax[i] += fx; ay[i] += fy; az[i] += fz;
                            ax[j] -= fx; ay[j] -= fy; az[j] -= fz;
                            d_epot = d_epot + dU_r;
                            dd_epot = dd_epot + (26.0 * inv_r12 - 7.0 * inv_r6);
                        }
                    }
                }
    }
    epot = 4.0 * epot;
    for (int32_t i = 0; i < n; ++i) { ax[i] *= 24.0; ay[i] *= 24.0; az[i] *= 24.0; }
    d_epot = 24.0 * d_epot;
    dd_epot = 24.0 * dd_epot;
    double te, td, tdd;
    orc_tail_corrections(p, &te, &td, &tdd);
    *epot_out = epot + te;
    *d_epot_out = d_epot + td;
    *dd_epot_out = dd_epot + tdd;
    if (n_candidates) *n_candidates = ncand;
    if (n_in_cutoff) *n_in_cutoff = ncut;
    free(cell); free(start); free(order);
    return ORC_OK;
}

int orc_verlet_step_cells(const orc_params *p, double *rx, double *ry, double *rz, double *vx,
                          double *vy, double *vz, double *ax, double *ay, double *az,
                          double *epot, double *ekin, double *d_epot, double *dd_epot)
{
    if (!p) return ORC_ERR_NULL;
    if (p->n <= 0) return ORC_ERR_N;
    orc_drift_wrap_halfkick(p, rx, ry, rz, vx, vy, vz, ax, ay, az);
    int rc_ = orc_compute_lj_cells(p, rx, ry, rz, ax, ay, az, epot, d_epot, dd_epot, NULL, NULL);
    if (rc_) return rc_;
    *ekin = orc_halfkick_ekin(p, vx, vy, vz, ax, ay, az);
    return ORC_OK;
}


/* ---------------------------------------------------------------------------------------
 * Consumers of the per-step reductions.
 * stats/md_means.f90:192-275 (md_means_add_sample, virial always present as in
 * md_simulation_program.f90:366-371), :284-366 (md_means_get), stats/stats_math.f90:56-61,
 * physics/thermodynamic_coefs.f90:87-205.
 * Quantity order: U, K, E, T, P, 1/K, dU, ddU, dU/K, dU^2/K, ddU/K.
 * ------------------------------------------------------------------------------------- */
void orc_means_init(orc_means *m, int32_t n, double volume)
{
    memset(m, 0, sizeof(*m));
    m->n = n;
    m->volume = volume;
}

int orc_means_add(orc_means *m, double epot, double ekin, double virial, double d_epot,
                  double dd_epot, double *temp_inst, double *press_inst)
{
    if (m->n <= 0) return ORC_ERR_N;
    if (!(m->volume > 0.0)) return ORC_ERR_VOLUME;
    const double npd = (double)m->n;
    const double rho = npd / m->volume;
    const double U = epot, K = ekin;
    const double E = U + K;
    const double T = 2.0 * K / (3.0 * npd);                    /* :221 */
    const double P = rho * T + virial / (3.0 * m->volume);     /* :227 */
    if (K <= 0.0) return ORC_ERR_EKIN;                         /* :243 */
    const double Kinv = 1.0 / K;
    const double dU = d_epot, ddU = dd_epot;
    const double dU_Kinv = dU * Kinv, ddU_Kinv = ddU * Kinv, dU2_Kinv = (dU * dU) * Kinv;
    const double x[11] = {U, K, E, T, P, Kinv, dU, ddU, dU_Kinv, dU2_Kinv, ddU_Kinv};
    m->n_samples += 1;
    for (int q = 0; q < 11; ++q) {
        m->sum[q] = m->sum[q] + x[q];
        m->sum_sq[q] = m->sum_sq[q] + x[q] * x[q];
    }
    if (temp_inst) *temp_inst = T;
    if (press_inst) *press_inst = P;
    return ORC_OK;
}

int orc_means_get(const orc_means *m, double *mean, double *std)
{
    if (m->n_samples <= 0) return ORC_ERR_N;
    const double inv_ns = 1.0 / (double)m->n_samples;
    for (int q = 0; q < 11; ++q) {
        mean[q] = m->sum[q] * inv_ns;
        const double m2 = m->sum_sq[q] * inv_ns;
        const double var = m2 - mean[q] * mean[q];             /* stats_math.f90:60 */
        std[q] = sqrt(var > 0.0 ? var : 0.0);
    }
    return ORC_OK;
}

/* out[16] = temperature, pressure, Ca_v, Ce_v, gamma, K_S_aux, K_S, K_S_inv, K_T, K_T_inv, Ca_p,
 * Ce_p, alpha_E1, alpha_E2, alpha_S, alpha_P.  Returns ORC_ERR_EKIN where the reference stops on a
 * vanishing denominator. */
int orc_thermo_coefs(const orc_params *p, double ekin_mean, double press_mean,
                     double ekinInv_mean, double d_epot_mean, double dd_epot_mean,
                     double d_epot_ekinInv_mean, double d_epot2_ekinInv_mean, double *out)
{
    const double npd = (double)p->n;
    const double f = 3.0 * npd - 3.0;                          /* :105 */
    if (f <= 0.0) return ORC_ERR_N;
    const double aux1 = 1.0 - 2.0 / f, aux2 = f / 2.0 - 1.0;   /* :109-110 */
    const double V = p->volume;
    const double T = 2.0 * ekin_mean / f;                      /* :118 */
    const double P = press_mean;
    double denom = 1.0 - aux1 * ekin_mean * ekinInv_mean;      /* :128 */
    if (fabs(denom) < 1.0e-14) return ORC_ERR_EKIN;
    const double Ca_v = 1.0 / denom;
    const double Ce_v = Ca_v / npd;
    if (fabs(Ce_v) < 1.0e-14) return ORC_ERR_EKIN;
    const double gamma = 1.0 / Ce_v + (aux2 / 3.0) * (d_epot_mean * ekinInv_mean - d_epot_ekinInv_mean); /* :137 */
    const double K_S_aux = ((npd * T * (1.0 + 2.0 * gamma - 1.0 / Ce_v)) / V) +
                           (dd_epot_mean - 2.0 * d_epot_mean) / (9.0 * V);                                /* :163 */
    const double K_S = K_S_aux - (aux2 * (d_epot2_ekinInv_mean - 2.0 * d_epot_mean * d_epot_ekinInv_mean +
                                          (d_epot_mean * d_epot_mean) * ekinInv_mean)) / (9.0 * V * V);  /* :164 */
    if (fabs(K_S) < 1.0e-14) return ORC_ERR_EKIN;
    const double K_S_inv = 1.0 / K_S;
    const double K_T = K_S - (T * Ca_v * (gamma * gamma)) / V; /* :172 */
    if (fabs(K_T) < 1.0e-14) return ORC_ERR_EKIN;
    const double K_T_inv = 1.0 / K_T;
    const double Ca_p = Ca_v * (K_S / K_T);
    const double Ce_p = Ca_p / npd;
    denom = (P * V / Ca_v) - (gamma * T);                      /* :188 */
    if (fabs(denom) < 1.0e-14) return ORC_ERR_EKIN;
    const double alpha_E1 = 1.0 / denom;
    denom = (1.0 / 3.0) * (aux1 * ekin_mean * d_epot_ekinInv_mean - d_epot_mean); /* :193 */
    if (fabs(denom) < 1.0e-14) return ORC_ERR_EKIN;
    const double alpha_E2 = 1.0 / denom;
    denom = gamma * T;
    if (fabs(denom) < 1.0e-14) return ORC_ERR_EKIN;
    const double alpha_S = -1.0 / denom;
    const double alpha_P = (Ca_v * gamma) / V * K_T_inv;       /* :203 */
    const double r[16] = {T, P, Ca_v, Ce_v, gamma, K_S_aux, K_S, K_S_inv, K_T, K_T_inv,
                          Ca_p, Ce_p, alpha_E1, alpha_E2, alpha_S, alpha_P};
    for (int q = 0; q < 16; ++q) out[q] = r[q];
    return ORC_OK;
}
