// lj_pair.cuh -- the per-pair arithmetic shared by every force kernel.
//
// Compiled with -fmad=false.  What is bit-identical to the reference loop body
// (/root/reference/scripts/physics/lj_potential_energy.f90:119-178) and what is not:
//   * mic(), dist2(): the reference's operations, operand order and roundings -- r2 and therefore the
//     cutoff decision r2 < rc2 (:132) are the reference's, for every pair, in every kernel.
//   * pair_terms() (full-matrix cross-check kernel, plain cell list): IEEE division and the reference's
//     operation order -- every summed term is bit-identical; only the ORDER of the sums differs.
//   * the default kernels (symmetric all-pairs tiles, neighbour-list walk) do NOT use pair_terms(): they
//     take 1/r2 from MUFU.RCP64H + one cubic step (<= 1 ulp, not correctly rounded), accumulate forces with
//     explicit FMAs and sum r^-12 and r^-6 separately (U, d_epot, dd_epot are formed from the two sums at
//     the end).  Their per-pair terms agree with the reference's to ~1 ulp, the results to ~1e-15 relative;
//     the tests hold them to 1e-12 (forces per atom against the sum of the pair-force magnitudes).
// The checker is a CPU restatement of the reference by the same author (no Fortran compiler here: parity unpinned).
#pragma once
#include <cuda_runtime.h>

namespace ljmd {

// Minimum image, physics/geometry_pbc.f90:86:  dx - L*dnint(dx*invL).
// rint (ties-to-even) differs from dnint (ties-away) only when |dx*invL| is exactly k+0.5,
// where |dx_mic| = L/2 > rc either way (rc < L/2 is enforced, md_types.f90:152), i.e. never
// for a pair inside the cutoff.  The FMA is exact-equivalent to the reference's separate
// multiply and subtract whenever L*k is representable (always for |k| <= 2, i.e. for any
// positions within two boxes of each other; wrapped positions give |k| <= 1).
//
// LJMD_MIC_SELECT (build-time experiment for A/B libraries, `make variant DEFS=-DLJMD_MIC_SELECT=1`;
// default 0): for WRAPPED positions |dx| <= L, so the image shift is 0 or -+L and dx -+ L is the very
// rounding of the reference's dx - L*(+-1).  The decision |dx| > L/2 can differ from
// dnint(dx*invL) only where |dx_mic| is within an ulp of L/2 > rc, i.e. for pairs outside the
// cutoff.  Same FP64-pipe instruction count (=1: DSETP compare) or two fewer per pair component
// (=2: 64-bit integer compare), and no FRND.F64 (9 cycles each on the pipe MUFU.RCP64H also uses).
// NOT valid for unwrapped input (|dx| > 1.5 L): never the default.
#ifndef LJMD_MIC_SELECT
#define LJMD_MIC_SELECT 0
#endif
__device__ __forceinline__ double mic(double dx, double L, double invL)
{
#if LJMD_MIC_SELECT == 0
    return fma(-L, rint(dx * invL), dx);
#else
    (void)invL;
    // sign handling on the bit patterns: fabs / negation / copysign of doubles would be emitted as
    // DADD-with-modifier moves on the FP64 pipe
    const int hi = __double2hiint(dx), lo = __double2loint(dx);
    const double neg_l = -L;
    const double sh = __hiloint2double(__double2hiint(neg_l) ^ (hi & (int)0x80000000), __double2loint(neg_l));
    const double shifted = __dadd_rn(dx, sh); // dx - L for dx > 0, dx + L for dx < 0
    const double half = 0.5 * L;              // exact
#if LJMD_MIC_SELECT == 1
    const bool far = fabs(dx) > half;
#else
    // non-negative doubles order like their bit patterns
    const unsigned long long mag = ((unsigned long long)((unsigned)hi & 0x7fffffffu) << 32) | (unsigned)lo;
    const bool far = mag > (unsigned long long)__double_as_longlong(half);
#endif
    return __hiloint2double(far ? __double2hiint(shifted) : hi, far ? __double2loint(shifted) : lo);
#endif
}

// Squared distance in the reference's order (:129): (dx*dx + dy*dy) + dz*dz, no contraction,
// so the strict cutoff test r2 < rc2 (:132) takes the same decision as the reference.
__device__ __forceinline__ double dist2(double dx, double dy, double dz)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

struct PairTerms {
    double fx, fy, fz; // force on i (without the factor 24), :148-150
    double u;          // inv_r12 - inv_r6, :140
    double du;         // -2*inv_r12 + inv_r6, :143
    double ddu;        // 26*inv_r12 - 7*inv_r6, :178
};

__device__ __forceinline__ PairTerms pair_terms(double dx, double dy, double dz, double r2)
{
    PairTerms t;
    const double inv_r2 = 1.0 / r2;                                       // :135 (IEEE division)
    const double inv_r6 = __dmul_rn(__dmul_rn(inv_r2, inv_r2), inv_r2);   // :136
    const double inv_r12 = __dmul_rn(inv_r6, inv_r6);                     // :137
    t.u = __dadd_rn(inv_r12, -inv_r6);                                    // :140
    t.du = __dadd_rn(__dmul_rn(-2.0, inv_r12), inv_r6);                   // :143
    const double m = -t.du;
    t.fx = __dmul_rn(__dmul_rn(m, dx), inv_r2);                           // :148
    t.fy = __dmul_rn(__dmul_rn(m, dy), inv_r2);                           // :149
    t.fz = __dmul_rn(__dmul_rn(m, dz), inv_r2);                           // :150
    t.ddu = __dadd_rn(__dmul_rn(26.0, inv_r12), -__dmul_rn(7.0, inv_r6)); // :178
    return t;
}

} // namespace ljmd
