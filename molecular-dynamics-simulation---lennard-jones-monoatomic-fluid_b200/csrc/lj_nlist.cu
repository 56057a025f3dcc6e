// lj_nlist.cu -- large-N path: counting-sort cells + Verlet neighbour list + ONE fused kernel per
// velocity-Verlet step (force, both half kicks, E_kin, next drift + wrap, unwrapped update,
// skin check, deterministic reductions).
//
// No reference counterpart: the reference is all-pairs only
// (/root/reference/scripts/physics/lj_potential_energy.f90:109-183).  The per-pair arithmetic,
// the minimum image (physics/geometry_pbc.f90:86), the strict cutoff r2 < rc2 evaluated on r2 in
// the reference's operation order (:129,:132) and the Verlet arithmetic (physics/verlet.f90:58-95)
// are the same as in the all-pairs path, so the set of interacting pairs is identical to the
// reference's and every per-pair term agrees to <= 1 ulp.
//
// Design
//   * The resident state lives in CELL-SORTED order (slot k holds original atom orig[k]); it is
//     re-sorted only when the list is rebuilt.  Contiguous slot ranges are spatial slabs.
//   * Rebuild (only when some atom moved more than skin/2 since the last build, decided on the
//     device, no host round trip): cell id of every atom -> histogram (integer atomics) ->
//     exclusive scan -> scatter -> per-cell sort by ORIGINAL index (= the stable counting sort,
//     bit-exact against the CPU definition) -> permute all state arrays -> full neighbour list
//     {j : |r_ij|_mic^2 < (rc+skin)^2}.
//     The list is stored in groups of four entries (one 16-byte load per thread), group g of all
//     slots contiguous (a warp streams 512 contiguous bytes per group from HBM, read once).
//     The build runs one CTA per work unit = (cell column, run of ~128 atoms along z) and stages
//     the unit's neighbourhood (9 columns x z-window) in shared memory; the periodic image of a
//     neighbour cell is a per-cell shift (exactly equal to the minimum image for every pair in
//     range).
//   * Step kernel: one thread per atom walks its list (ordered pairs: every thread accumulates
//     only its own atom => no atomics, fixed summation order, bit-reproducible), gathers the
//     packed positions (x, y, z, 0: one 256-bit load = one 32-byte sector per neighbour) from
//     L1/L2 and finishes the velocity-Verlet step for its atom in the epilogue.  Measured on
//     B200 (profiles/): the gathers are bound by the L1 tag stage (~32 clocks per scattered warp
//     gather); a variant that staged the neighbourhood in shared memory and kept 16-bit indices
//     was slower (staging re-reads every position ~12x from L2).  Warps whose atoms all sit in interior
//     cells skip the minimum image (dx - L*0 = dx exactly); boundary warps use the rint() form.
//     The last CTA reduces the scalar partials in a fixed order.
//   * Skin criterion (exact): a pair absent from the list had r > rc+skin at build time; while no
//     atom has moved more than skin/2 it is still beyond rc.  The epilogue measures the displacement
//     of the positions it just produced against the build positions.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ljmd_internal.h"
#include "lj_pair.cuh"
#include "reduce.cuh"

namespace ljmd {

#ifndef NL_BS_VALUE
#define NL_BS_VALUE 128
#endif
constexpr int NL_BS = NL_BS_VALUE; // threads per CTA (= target atoms per work unit) of the per-atom kernels

// Kernel argument block (passed by value).
struct NlArgs {
    SimConst sc;
    int nc, ncells;
    double scale;      // nc / L
    double rl2;        // (rc + skin)^2
    double half_skin2; // (skin/2)^2
    int maxnb;
    size_t stride;
    int nzg, nunits;   // z-groups per column, work units (= CTAs of the build / step kernels)
    int cap_stage;     // shared-memory staging capacity (atoms)
    double *rb[2];     // position buffers: x array; y, z follow at +n, +2n
    double4 *p4[2];    // the same positions packed (x, y, z, 0): one 32-byte sector per atom for the gathers
    double *v, *a;     // x arrays; y, z at +n, +2n
    double *ru;        // unwrapped x array or nullptr
    double *r0;        // [3][n] positions at the last list build
    double *alt;       // [narr][n] permutation scratch
    double *arr[15];   // the state arrays that are permuted at a rebuild
    int narr;
    int *cell_of_slot, *cell_count, *cell_start, *cell_cursor, *tile_sum;
    unsigned long long *keys;
    int *orig, *orig_alt, *slot_of_orig;
    unsigned char *interior;
    int *nb_cnt;
    int *nbr;          // the list: global slot indices, see nl_entry()
    NlCtl *ctl;
    StepCtl *sctl;
    double *spart; // [nblk][5]
    int nblk;
    double *obs;   // [4]
    double *mom;   // [3] total momentum of the atoms handled by the last step launch
    // ---- slab sharding over GPUs (sh_on): this rank owns the cell layers cx in [xlo, xhi)
    int sh_on, rank, world, xlo, xhi;
    double4 *peer_p4[LJMD_MAX_RANKS]; // p4 base ([2][n]) of every rank, mapped into this process
    int *peer_flags[LJMD_MAX_RANKS];  // flag block ([2][LJMD_MAX_RANKS] ints) of every rank
    int *flags;                       // this rank's flag block (the peers write into it)
    int *dbg_exec;                    // debug: CTAs that executed the body of sharded step seq (or nullptr)
    unsigned long long *dbg_time;     // LJMD_SLAB_TIMELINE: [1024][4] globaltimer of {first CTA start, last flag seen, last CTA end, publish}
    unsigned long long spin_budget_ns; // longest wait for the peers' flags before the run is declared dead
    // slab-LOCAL rebuild (sl_on): the sort kernels work on slot pieces instead of [0, n).  sl_rng (device):
    // [0..2] start, [3..5] length of the CANDIDATE pieces {left neighbour's last two cell layers, the own slab,
    // right neighbour's first two layers} under the OLD cell table, [6] their total; [8..10], [11..13], [14]: the
    // same for the TARGET pieces {left halo layer, own slab, right halo layer} under the NEW cell table.
    int sl_on;
    const int *sl_rng;
    int eval_sel;  // rebuild kernels: 0 = the list is for rb[cur], 1 = for rb[1-cur]
    unsigned long long *pair_counts;
};

// Entry m of slot i: the list is stored in groups of four entries (one 16-byte load per thread),
// group g of all slots contiguous (a warp reads 512 contiguous bytes per group).
__host__ __device__ __forceinline__ size_t nl_entry(size_t stride, int i, int m)
{
    return (((size_t)(m >> 2) * stride + (size_t)i) << 2) + (size_t)(m & 3);
}

// 256-bit read-only load (LDG.E.256 on sm_100a): one packed position.
__device__ __forceinline__ double4 nl_ld_pos(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void nl_st_pos(double4 *p, double x, double y, double z)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(0.0) : "memory");
}

// 1/x to <= 1 ulp (see lj_allpairs.cu).
__device__ __forceinline__ double nl_rcp(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0, 1.0);
    const double t = fma(e, e, e);
    return fma(y0, t, y0);
}

// Cell coordinate of x.  In-box x (0 <= x <= L, what wrap_positions produces,
// geometry_pbc.f90:53-57) bins as floor(x*scale) clamped to [0, nc-1]; out-of-box x (the force
// routine also accepts unwrapped input, geometry_pbc.f90:18-19) is wrapped for binning only.
__device__ __forceinline__ int nl_cell_coord(double x, double L, double invL, double scale, int nc, bool &outside)
{
    if (x < 0.0 || x > L) { x = x - L * floor(x * invL); outside = true; }
    int c = (int)floor(x * scale);
    c = max(c, 0);
    return min(c, nc - 1);
}

// Slab-local rebuild: t-th slot of the three pieces at rng[0..2] / rng[3..5]; -1 past the end.
__device__ __forceinline__ int nl_sl_slot(const int *rng, int t)
{
    if (t < rng[3]) return rng[0] + t;
    t -= rng[3];
    if (t < rng[4]) return rng[1] + t;
    t -= rng[4];
    return t < rng[5] ? rng[2] + t : -1;
}
// Cell layer (x index) of cell c relative to the owned layers: -1 = left halo layer, 0 .. = owned, nown = right
// halo layer, anything else = not this rank's business.  Periodic.
__device__ __forceinline__ int nl_sl_rel_layer(const NlArgs &a, int c)
{
    const int cx = c / (a.nc * a.nc);
    if (cx >= a.xlo && cx < a.xhi) return cx - a.xlo;
    if (cx == (a.xlo - 1 + a.nc) % a.nc) return -1;
    if (cx == a.xhi % a.nc) return a.xhi - a.xlo;
    return -2;
}

// ------------------------------------------------------------------------------------------
// Rebuild sequence.  Every kernel returns at once unless ctl->need_rebuild is set.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nl_bin_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    const int n = a.sc.n;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.sl_on) k = nl_sl_slot(a.sl_rng, k); // candidates of a slab-local rebuild
    if (k >= n || k < 0) return;
    const double *px = a.rb[a.ctl->cur ^ a.eval_sel];
    bool outside = false;
    const int cx = nl_cell_coord(px[k], a.sc.L, a.sc.invL, a.scale, a.nc, outside);
    const int cy = nl_cell_coord(px[(size_t)n + k], a.sc.L, a.sc.invL, a.scale, a.nc, outside);
    const int cz = nl_cell_coord(px[2 * (size_t)n + k], a.sc.L, a.sc.invL, a.scale, a.nc, outside);
    const int c = (cx * a.nc + cy) * a.nc + cz;
    a.cell_of_slot[k] = c;
    // integer histogram: order-independent.  Slab-local: every rank counts the cells of its OWN layers only (it
    // sees every atom that can end up there); the ranks' histograms are then added by one all-reduce.
    if (!a.sl_on || (cx >= a.xlo && cx < a.xhi)) atomicAdd(&a.cell_count[c], 1);
    if (outside) a.ctl->unwrapped = 1;
}

// Exclusive scan of the histogram in three small kernels: (a) every CTA scans a tile of 1024 cells
// and publishes the tile total, (b) one CTA scans the tile totals, (c) the tile offsets are added.
// (a) also re-arms the histogram and the scatter cursors.
__device__ __forceinline__ int nl_block_exscan(int v, int *wsum, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads(); // wsum may still be read from a previous call
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    total = wsum[31];
    return (warp > 0 ? wsum[warp - 1] : 0) + x - v;
}

__global__ void __launch_bounds__(1024) nl_scan_tiles_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    __shared__ int wsum[32];
    const int c = blockIdx.x * 1024 + threadIdx.x;
    const int v = c < a.ncells ? a.cell_count[c] : 0;
    int total;
    const int ex = nl_block_exscan(v, wsum, total);
    if (c < a.ncells) { a.cell_start[c] = ex; a.cell_cursor[c] = 0; a.cell_count[c] = 0; }
    if (threadIdx.x == 0) a.tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) nl_scan_sums_kernel(NlArgs a, int ntiles)
{
    if (!a.ctl->need_rebuild) return;
    __shared__ int wsum[32];
    int carry = 0;
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int v = t < ntiles ? a.tile_sum[t] : 0;
        int total;
        const int ex = nl_block_exscan(v, wsum, total);
        if (t < ntiles) a.tile_sum[t] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) a.cell_start[a.ncells] = carry;
}

__global__ void __launch_bounds__(1024) nl_scan_add_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    const int c = blockIdx.x * 1024 + threadIdx.x;
    if (c < a.ncells) a.cell_start[c] += a.tile_sum[blockIdx.x];
}

__global__ void __launch_bounds__(256) nl_fill_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.sl_on) k = nl_sl_slot(a.sl_rng, k); // candidates (pieces saved before the scan replaced the cell table)
    if (k >= a.sc.n || k < 0) return;
    const int c = a.cell_of_slot[k];
    if (a.sl_on && nl_sl_rel_layer(a, c) == -2) return; // ends up in a cell this rank neither owns nor sees
    const int pos = a.cell_start[c] + atomicAdd(&a.cell_cursor[c], 1);
    a.keys[pos] = ((unsigned long long)(unsigned)a.orig[k] << 32) | (unsigned)k;
}

// One thread per cell: sort the (short) segment by original index => stable counting sort.
__global__ void __launch_bounds__(128) nl_sortcell_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.sl_on) { // the cells of the left halo layer, the owned layers and the right halo layer
        const int layer = a.nc * a.nc, nl = a.xhi - a.xlo + 2;
        if (c >= nl * layer) return;
        const int li = c / layer - 1; // -1 .. nown
        const int cx = li < 0 ? (a.xlo - 1 + a.nc) % a.nc : (a.xlo + li) % a.nc;
        c = cx * layer + c % layer;
    }
    if (c >= a.ncells) return;
    const int b = a.cell_start[c], e = a.cell_start[c + 1];
    for (int k = b + 1; k < e; ++k) {
        const unsigned long long v = a.keys[k];
        int m = k - 1;
        while (m >= b && a.keys[m] > v) { a.keys[m + 1] = a.keys[m]; --m; }
        a.keys[m + 1] = v;
    }
}

// New slot k takes the atom that sat in old slot (keys[k] & 0xffffffff).
__global__ void __launch_bounds__(256) nl_permute_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    const int n = a.sc.n;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.sl_on) k = nl_sl_slot(a.sl_rng + 8, k); // target slots: halo layer | own slab | halo layer, new cell table
    if (k >= n || k < 0) return;
    const unsigned long long key = a.keys[k];
    const int old = (int)(key & 0xffffffffu), o = (int)(key >> 32);
    for (int q = 0; q < a.narr; ++q) a.alt[(size_t)q * n + k] = a.arr[q][old];
    a.orig_alt[k] = o;
    a.slot_of_orig[o] = k;
}

__global__ void __launch_bounds__(256) nl_copyback_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    const int n = a.sc.n;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.sl_on) k = nl_sl_slot(a.sl_rng + 8, k);
    if (k >= n || k < 0) return;
    for (int q = 0; q < a.narr; ++q) a.arr[q][k] = a.alt[(size_t)q * n + k];
    a.orig[k] = a.orig_alt[k];
    // packed copy of the positions the list is built for (arrays 0-2 are rb[0], 3-5 rb[1])
    const int ev = a.ctl->cur ^ a.eval_sel;
    nl_st_pos(a.p4[ev] + k, a.alt[(size_t)(3 * ev) * n + k], a.alt[(size_t)(3 * ev + 1) * n + k],
              a.alt[(size_t)(3 * ev + 2) * n + k]);
}

// Slab-local rebuild: slot pieces from the cell table.  which = 0 (before the scan, OLD table): candidates =
// {two layers left of the slab, the slab, two layers right of it}; which = 1 (after the scan, NEW table):
// targets = {one layer, the slab, one layer}.  With >= 2 layers per rank every piece is contiguous.
__global__ void nl_sl_ranges_kernel(NlArgs a, int *rng, int which)
{
    if (threadIdx.x != 0 || !a.ctl->need_rebuild) return;
    const int layer = a.nc * a.nc, h = which ? 1 : 2;
    const int l0 = ((a.xlo - h) % a.nc + a.nc) % a.nc, l2 = a.xhi % a.nc;
    int *r = rng + 8 * which;
    r[0] = a.cell_start[l0 * layer]; r[3] = a.cell_start[(l0 + h) * layer] - r[0];
    r[1] = a.cell_start[a.xlo * layer]; r[4] = a.cell_start[a.xhi * layer] - r[1];
    r[2] = a.cell_start[l2 * layer]; r[5] = a.cell_start[(l2 + h) * layer] - r[2];
    r[6] = r[3] + r[4] + r[5];
}

// Two cell layers (first cell layer l0, periodic) of every permuted state array + the original indices <->
// a dense buffer [narr + 1][cap]: what a rank sends to / receives from a neighbour at a slab-local rebuild.
__global__ void __launch_bounds__(256) nl_halo_pack_kernel(NlArgs a, double *buf, int cap, int l0, int unpack)
{
    const int layer = a.nc * a.nc;
    const int sb = a.cell_start[l0 * layer], cnt = min(cap, a.cell_start[(l0 + 2) * layer] - sb);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    if (!unpack) {
        for (int q = 0; q < a.narr; ++q) buf[(size_t)q * cap + t] = a.arr[q][sb + t];
        buf[(size_t)a.narr * cap + t] = (double)a.orig[sb + t];
    } else {
        for (int q = 0; q < a.narr; ++q) a.arr[q][sb + t] = buf[(size_t)q * cap + t];
        a.orig[sb + t] = (int)buf[(size_t)a.narr * cap + t];
    }
}

// ------------------------------------------------------------------------------------------
// Work units.  Unit u = (column col = u / nzg, z-group g = u % nzg) owns the cells
// z0 = g*nc/nzg .. z1 = (g+1)*nc/nzg - 1 of its column: one contiguous slot range [ib, ie).  Its
// neighbourhood is, for each of the 9 neighbour columns, the cells z0-1 .. z1+1 (periodic): one or
// two contiguous slot runs.  Run r = 2*(column index) + piece is staged at shared-memory offset
// off[r]; the same table is rebuilt from cell_start by the build kernel and by every step.
// ------------------------------------------------------------------------------------------
constexpr int NL_RUNS = 18;
struct UnitTable {
    int sb[NL_RUNS];      // first slot of the run
    int off[NL_RUNS + 1]; // shared-memory offset of the run (prefix sum of the run lengths)
    int za[NL_RUNS], zb[NL_RUNS]; // the run covers cells [za, zb) of its column
    int ib, ie, cx, cy, z0, z1;
};

__device__ __forceinline__ void nl_unit_setup(const NlArgs &a, int unit, UnitTable &ut)
{
    const int nc = a.nc;
    const int g = unit % a.nzg, col = unit / a.nzg;
    const int cy = col % nc, cx = col / nc;
    const int z0 = g * nc / a.nzg, z1 = (g + 1) * nc / a.nzg - 1;
    __shared__ int len[NL_RUNS];
    if (threadIdx.x < NL_RUNS) {
        const int r = threadIdx.x, cidx = r >> 1, piece = r & 1;
        const int nx = (cx + cidx / 3 - 1 + nc) % nc, ny = (cy + cidx % 3 - 1 + nc) % nc;
        const int nb = (nx * nc + ny) * nc;
        const int zlo = z0 - 1, zhi = z1 + 1;
        int za = 0, zb = 0;
        if (zhi - zlo + 1 >= nc) { if (!piece) { za = 0; zb = nc; } }              // whole column
        else if (zlo < 0) { if (!piece) { za = nc - 1; zb = nc; } else { za = 0; zb = zhi + 1; } }
        else if (zhi >= nc) { if (!piece) { za = zlo; zb = nc; } else { za = 0; zb = 1; } }
        else { if (!piece) { za = zlo; zb = zhi + 1; } }
        ut.za[r] = za; ut.zb[r] = zb;
        const int s0 = a.cell_start[nb + za];
        ut.sb[r] = s0;
        len[r] = a.cell_start[nb + zb] - s0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int r = 0; r < NL_RUNS; ++r) { ut.off[r] = o; o += len[r]; }
        ut.off[NL_RUNS] = o;
        const int cb = (cx * nc + cy) * nc;
        ut.ib = a.cell_start[cb + z0];
        ut.ie = a.cell_start[cb + z1 + 1];
        ut.cx = cx; ut.cy = cy; ut.z0 = z0; ut.z1 = z1;
    }
    __syncthreads();
}

// Stage the neighbourhood positions: sx/sy/sz[q] for q in [0, off[NL_RUNS]).  Four independent
// 256-bit loads per thread are in flight before the first shared-memory store.
__device__ __forceinline__ void nl_stage(const UnitTable &ut, const double4 *__restrict__ pos, double *sx, double *sy,
                                         double *sz, int total)
{
    constexpr int U = 4;
    const int nt = blockDim.x;
    int r = 0;
    for (int q0 = threadIdx.x; q0 < total; q0 += U * nt) {
        double4 p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = q0 + u * nt;
            if (q < total) {
                while (q >= ut.off[r + 1]) ++r;
                p[u] = nl_ld_pos(pos + (ut.sb[r] + q - ut.off[r]));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = q0 + u * nt;
            if (q < total) { sx[q] = p[u].x; sy[q] = p[u].y; sz[q] = p[u].z; }
        }
    }
}

// Candidate test of the build: r2 in the reference's operation order (lj_potential_energy.f90:
// 119-129).  With SHIFT the minimum image is the per-cell image shift k (dx - L*k, k in {-1,0,1},
// the same single rounding as geometry_pbc.f90:86 for every pair in range); without it (unwrapped
// input) the rint() form is used.
template <bool SHIFT>
__device__ __forceinline__ double nl_build_r2(double xi, double yi, double zi, double xj, double yj, double zj,
                                              double sx_, double sy_, double sz_, double L, double invL)
{
    double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    if (SHIFT) { dx = dx - sx_; dy = dy - sy_; dz = dz - sz_; }
    else { dx = mic(dx, L, invL); dy = mic(dy, L, invL); dz = mic(dz, L, invL); }
    return dist2(dx, dy, dz);
}

template <bool STAGED, bool SHIFT>
__device__ __forceinline__ int nl_build_atom(const NlArgs &a, const UnitTable &ut, const double4 *__restrict__ pos,
                                             const double *sx, const double *sy, const double *sz, int i, double xi,
                                             double yi, double zi, int cz)
{
    const int nc = a.nc;
    const double L = a.sc.L, invL = a.sc.invL, rl2 = a.rl2;
    int cnt = 0;
    for (int cidx = 0; cidx < 9; ++cidx) {
        const int ux = ut.cx + cidx / 3 - 1, uy = ut.cy + cidx % 3 - 1;
        const int nx = (ux + nc) % nc, ny = (uy + nc) % nc;
        const double shx = (ux < 0) ? -L : (ux >= nc) ? L : 0.0; // L * k, exact
        const double shy = (uy < 0) ? -L : (uy >= nc) ? L : 0.0;
        const int nb = (nx * nc + ny) * nc;
        for (int w = cz - 1; w <= cz + 1; ++w) {
            const int cw = (w + nc) % nc;
            const double shz = (w < 0) ? -L : (w >= nc) ? L : 0.0;
            const int s0 = a.cell_start[nb + cw], s1 = a.cell_start[nb + cw + 1];
            int q0 = 0;
            if (STAGED) {
                const int r = 2 * cidx + ((cw >= ut.za[2 * cidx] && cw < ut.zb[2 * cidx]) ? 0 : 1);
                q0 = ut.off[r] + (s0 - ut.sb[r]);
            }
            for (int k = 0; k < s1 - s0; ++k) {
                double xj, yj, zj;
                if (STAGED) { xj = sx[q0 + k]; yj = sy[q0 + k]; zj = sz[q0 + k]; }
                else { const double4 pj = nl_ld_pos(pos + s0 + k); xj = pj.x; yj = pj.y; zj = pj.z; }
                const double r2 = nl_build_r2<SHIFT>(xi, yi, zi, xj, yj, zj, shx, shy, shz, L, invL);
                if (r2 < rl2 && (s0 + k) != i) {
                    if (cnt < a.maxnb) a.nbr[nl_entry(a.stride, i, cnt)] = s0 + k;
                    ++cnt;
                }
            }
        }
    }
    return cnt;
}

// Full neighbour list of the unit's atoms.  Membership is decided on r2 computed exactly like the
// reference computes it.  The unit's neighbourhood is staged in shared memory when it fits (it
// does for any roughly homogeneous system); otherwise the candidates are read from global memory.
__global__ void __launch_bounds__(NL_BS) nl_build_kernel(NlArgs a)
{
    if (!a.ctl->need_rebuild) return;
    extern __shared__ double smem[];
    __shared__ UnitTable ut;
    const int n = a.sc.n, nc = a.nc;
    // sharded: only the units of the owned cell layers get lists (the sort above covered all atoms)
    const int ucx = (int)(blockIdx.x / a.nzg) / nc;
    const bool mine = !a.sh_on || (ucx >= a.xlo && ucx < a.xhi);
    if (mine) nl_unit_setup(a, blockIdx.x, ut);
    const int total = mine ? ut.off[NL_RUNS] : 0;
    const int ev = a.ctl->cur ^ a.eval_sel;
    const double4 *pos = a.p4[ev];
    double *sx = smem, *sy = sx + a.cap_stage, *sz = sy + a.cap_stage;
    const bool staged = total <= a.cap_stage;
    const bool shift = a.ctl->unwrapped == 0;
    if (mine && staged) nl_stage(ut, pos, sx, sy, sz, total);
    __syncthreads();
    const double L = a.sc.L, invL = a.sc.invL;
    for (int i = (mine ? ut.ib : 0) + threadIdx.x; i < (mine ? ut.ie : 0); i += NL_BS) {
        const double4 pi = nl_ld_pos(pos + i);
        const double xi = pi.x, yi = pi.y, zi = pi.z;
        bool outside = false;
        const int cz = nl_cell_coord(zi, L, invL, a.scale, nc, outside);
        int cnt;
        if (staged) {
            if (shift) cnt = nl_build_atom<true, true>(a, ut, pos, sx, sy, sz, i, xi, yi, zi, cz);
            else cnt = nl_build_atom<true, false>(a, ut, pos, sx, sy, sz, i, xi, yi, zi, cz);
        } else {
            if (shift) cnt = nl_build_atom<false, true>(a, ut, pos, sx, sy, sz, i, xi, yi, zi, cz);
            else cnt = nl_build_atom<false, false>(a, ut, pos, sx, sy, sz, i, xi, yi, zi, cz);
        }
        if (cnt > a.maxnb) atomicMax(&a.ctl->overflow, cnt);
        cnt = min(cnt, a.maxnb);
        a.nb_cnt[i] = cnt;
        for (int m = cnt; m & 3; ++m) a.nbr[nl_entry(a.stride, i, m)] = i; // pad the last group (masked in the walk)
        a.r0[i] = xi; a.r0[(size_t)n + i] = yi; a.r0[2 * (size_t)n + i] = zi;
        const int cx = ut.cx, cy = ut.cy;
        const bool inter = cx >= 1 && cx <= nc - 2 && cy >= 1 && cy <= nc - 2 && cz >= 1 && cz <= nc - 2;
        a.interior[i] = inter ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&a.ctl->ticket_build, 1u);
        if (t == gridDim.x - 1) { // every CTA has passed its need_rebuild test
            a.ctl->ticket_build = 0;
            a.ctl->maxd2_bits = 0ull;
            a.ctl->rebuilds += 1;
            __threadfence();
            a.ctl->need_rebuild = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Per-pair body, branch-free: an entry outside the cutoff contributes exact zeros.
// ------------------------------------------------------------------------------------------
template <bool MIC, bool MASKED>
__device__ __forceinline__ void nl_pair(double xi, double yi, double zi, const double4 pj, bool keep, double L,
                                        double invL, double rc2, double &fx, double &fy, double &fz, double &s12,
                                        double &s6)
{
    double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z; // lj_potential_energy.f90:119-121
    if (MIC) { dx = mic(dx, L, invL); dy = mic(dy, L, invL); dz = mic(dz, L, invL); } // :124-126
    const double r2 = dist2(dx, dy, dz);             // :129
    bool in = r2 < rc2;                              // :132
    if (MASKED) in = in && keep;                     // padding entries of the last group
    double ir2 = nl_rcp(r2);
    ir2 = in ? ir2 : 0.0;
    const double ir6 = __dmul_rn(__dmul_rn(ir2, ir2), ir2);
    const double ir12 = __dmul_rn(ir6, ir6);
    s12 = __dadd_rn(s12, ir12);
    s6 = __dadd_rn(s6, ir6);
    const double g = __dmul_rn(fma(2.0, ir12, -ir6), ir2); // -dU_r / r^2 (:143,148)
    fx = fma(g, dx, fx); fy = fma(g, dy, fy); fz = fma(g, dz, fz);
}

// Walk of one atom's list.  The index groups stream from HBM (read once: cache-streaming loads, a
// ring of NL_PF groups is kept in flight); the packed positions are gathered from L1/L2.
constexpr int NL_PF = 3;
constexpr int NL_NS = 8; // per-CTA partial sums: S12, S6, sum vx^2, vy^2, vz^2, sum vx, vy, vz

// TPA lanes share one atom: lane `sub` takes the entry groups sub, sub+TPA, ... (TPA = 4 gives small
// systems -- and the slabs of a sharded run -- four times as many threads; the lane partials are
// combined afterwards in a fixed butterfly order).
template <bool MIC, int TPA>
__device__ __forceinline__ void nl_walk(const int4 *__restrict__ nb, size_t stride, int cnt, int sub,
                                        const double4 *__restrict__ pos, double xi, double yi, double zi, double L,
                                        double invL, double rc2, double &fx, double &fy, double &fz, double &s12,
                                        double &s6)
{
    const int ngrp = (cnt + 3) >> 2, nfull = cnt >> 2;
    int4 ring[NL_PF];
#pragma unroll
    for (int u = 0; u < NL_PF; ++u) {
        ring[u] = make_int4(0, 0, 0, 0);
        if (sub + u * TPA < ngrp) ring[u] = __ldcs(nb + (size_t)(sub + u * TPA) * stride);
    }
    int g = sub;
    for (; g < nfull; g += TPA) {
        const int4 j = ring[0];
#pragma unroll
        for (int u = 0; u + 1 < NL_PF; ++u) ring[u] = ring[u + 1];
        if (g + NL_PF * TPA < ngrp) ring[NL_PF - 1] = __ldcs(nb + (size_t)(g + NL_PF * TPA) * stride);
        const double4 pa = nl_ld_pos(pos + j.x), pb = nl_ld_pos(pos + j.y), pc = nl_ld_pos(pos + j.z),
                      pd = nl_ld_pos(pos + j.w);
        nl_pair<MIC, false>(xi, yi, zi, pa, true, L, invL, rc2, fx, fy, fz, s12, s6);
        nl_pair<MIC, false>(xi, yi, zi, pb, true, L, invL, rc2, fx, fy, fz, s12, s6);
        nl_pair<MIC, false>(xi, yi, zi, pc, true, L, invL, rc2, fx, fy, fz, s12, s6);
        nl_pair<MIC, false>(xi, yi, zi, pd, true, L, invL, rc2, fx, fy, fz, s12, s6);
    }
    if (g < ngrp) { // this lane owns the last, partly filled group: padding entries point at the atom itself and are masked
        const int4 j = ring[0];
        const int rem = cnt & 3;
        const double4 pa = nl_ld_pos(pos + j.x), pb = nl_ld_pos(pos + j.y), pc = nl_ld_pos(pos + j.z);
        nl_pair<MIC, true>(xi, yi, zi, pa, rem > 0, L, invL, rc2, fx, fy, fz, s12, s6);
        nl_pair<MIC, true>(xi, yi, zi, pb, rem > 1, L, invL, rc2, fx, fy, fz, s12, s6);
        nl_pair<MIC, true>(xi, yi, zi, pc, rem > 2, L, invL, rc2, fx, fy, fz, s12, s6);
    }
}

// The same walk with the position gathers of the NEXT group issued before the current group is
// evaluated (software pipelining; identical arithmetic and summation order).  Small systems have
// few warps per SM and are bound by the gather round trip (~2000 clocks per group at N=32k,
// profiles/r1_ncu_nlist_step_cfg3.txt), so they can afford the extra registers.
template <bool MIC>
__device__ __forceinline__ void nl_walk_pipe(const int4 *__restrict__ nb, size_t stride, int cnt,
                                             const double4 *__restrict__ pos, double xi, double yi, double zi,
                                             double L, double invL, double rc2, double &fx, double &fy, double &fz,
                                             double &s12, double &s6)
{
    const int ngrp = (cnt + 3) >> 2, nfull = cnt >> 2;
    if (ngrp == 0) return;
    const int4 zero = make_int4(0, 0, 0, 0);
    int4 j = __ldcs(nb);
    int4 jn1 = (1 < ngrp) ? __ldcs(nb + stride) : zero;
    int4 jn2 = (2 < ngrp) ? __ldcs(nb + 2 * stride) : zero;
    double4 pa = nl_ld_pos(pos + j.x), pb = nl_ld_pos(pos + j.y), pc = nl_ld_pos(pos + j.z), pd = nl_ld_pos(pos + j.w);
    for (int g = 0;; ++g) {
        const bool has_next = g + 1 < ngrp;
        double4 qa = pa, qb = pb, qc = pc, qd = pd;
        int4 jn3 = zero;
        if (has_next) { // every entry of a group is a valid slot (padding = the atom itself)
            qa = nl_ld_pos(pos + jn1.x); qb = nl_ld_pos(pos + jn1.y); qc = nl_ld_pos(pos + jn1.z); qd = nl_ld_pos(pos + jn1.w);
            if (g + 3 < ngrp) jn3 = __ldcs(nb + (size_t)(g + 3) * stride);
        }
        if (g < nfull) {
            nl_pair<MIC, false>(xi, yi, zi, pa, true, L, invL, rc2, fx, fy, fz, s12, s6);
            nl_pair<MIC, false>(xi, yi, zi, pb, true, L, invL, rc2, fx, fy, fz, s12, s6);
            nl_pair<MIC, false>(xi, yi, zi, pc, true, L, invL, rc2, fx, fy, fz, s12, s6);
            nl_pair<MIC, false>(xi, yi, zi, pd, true, L, invL, rc2, fx, fy, fz, s12, s6);
        } else {
            const int rem = cnt & 3;
            nl_pair<MIC, true>(xi, yi, zi, pa, rem > 0, L, invL, rc2, fx, fy, fz, s12, s6);
            nl_pair<MIC, true>(xi, yi, zi, pb, rem > 1, L, invL, rc2, fx, fy, fz, s12, s6);
            nl_pair<MIC, true>(xi, yi, zi, pc, rem > 2, L, invL, rc2, fx, fy, fz, s12, s6);
        }
        if (!has_next) break;
        pa = qa; pb = qb; pc = qc; pd = qd;
        jn1 = jn2; jn2 = jn3;
    }
}

__device__ __forceinline__ double nl_drift_wrap(double x, double v, double acc, const SimConst &sc)
{
    const double xn = (x + v * sc.dt) + acc * sc.dt_square_half; // verlet.f90:58
    return xn - sc.L * floor(xn * sc.invL);                      // geometry_pbc.f90:54-56
}

// Squared displacement of (x,y,z) from the build position of slot i (minimum image).
__device__ __forceinline__ double nl_disp2(const NlArgs &a, int i, double x, double y, double z)
{
    const int n = a.sc.n;
    const double dx = mic(x - a.r0[i], a.sc.L, a.sc.invL);
    const double dy = mic(y - a.r0[(size_t)n + i], a.sc.L, a.sc.invL);
    const double dz = mic(z - a.r0[2 * (size_t)n + i], a.sc.L, a.sc.invL);
    return dx * dx + dy * dy + dz * dz;
}

// CTA-wide maximum of a non-negative double, folded into ctl->maxd2_bits (bit patterns of
// non-negative doubles order like unsigned integers => the atomic max is order-independent).
__device__ __forceinline__ void nl_block_max(double d2, double *scratch, NlCtl *ctl)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = d2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = scratch[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, scratch[w]);
        if (m > 0.0) atomicMax(&ctl->maxd2_bits, (unsigned long long)__double_as_longlong(m));
    }
}

// ------------------------------------------------------------------------------------------
// The step kernel.  MODE 0: forces + observables at rb[cur] (compute_lj_potential_energy).
// MODE 1: one verlet_step: forces at rb[1-cur] (already drifted), both half kicks, E_kin,
// unwrapped update, r(t+2dt) -> rb[cur], skin check; the last CTA flips cur.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nl_ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long nl_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int NL_ERR_FLAG_TIMEOUT = 0x20000000; // NlCtl::overflow code: a peer's per-step flag never arrived
constexpr int NL_ERR_OWN_BLOCKS = 0x40000000;   // NlCtl::overflow code: more owned atoms than the launch has CTAs for
constexpr int NL_FLAG_POISON = 0x7fffffff;      // flag value of a rank that failed: "every step done, skin violated"
__device__ __forceinline__ void nl_st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// SH (slab sharding, MODE 1 only): the kernel handles the slots of the owned cell layers.  Step s+1
// starts only when every rank has published its flag of step s (value = (s << 1) | violated, in
// the parity slot s & 1 of this rank's flag block): at that point the halo positions written by
// the neighbours into p4 are visible, and nobody still reads the buffer this step overwrites.  If
// any rank reported a skin violation the kernel halts (no state change) until the host has
// re-gathered the state and rebuilt the lists.
// Single-GPU variants: <= 64 registers (32 warps per SM hide the gather latency).  The sharded
// variants need a few more registers; they get up to 85 rather than being squeezed into 64 with
// local-memory spills.
// Resident threads per SM the register allocation is sized for (build-time knobs for A/B libraries,
// `make variant`): PIPE 512, TPA > 1 640, sharded / force-only 768, plain step 1024.
#ifndef LJMD_NL_THREADS_SH
#define LJMD_NL_THREADS_SH 1024 // 8 CTAs per SM: the 1050 CTAs of the fullest rank of an 8-GPU run at N = 2^20 are ONE wave
#endif
#ifndef LJMD_NL_PUBLISH_IN_KERNEL
#define LJMD_NL_PUBLISH_IN_KERNEL 1 // the last CTA of a sharded step publishes the flags itself (0: a 1-warp kernel after every step)
#endif
#ifndef LJMD_NL_THREADS_STEP
#define LJMD_NL_THREADS_STEP 1024
#endif
#ifndef LJMD_NL_THREADS_PIPE
#define LJMD_NL_THREADS_PIPE 512
#endif
template <int MODE, bool TRACK, bool SH, int TPA, bool PIPE = false>
__global__ void __launch_bounds__(NL_BS, PIPE ? LJMD_NL_THREADS_PIPE / NL_BS : TPA > 1 ? 640 / NL_BS
                                         : (SH || MODE == 0) ? LJMD_NL_THREADS_SH / NL_BS : LJMD_NL_THREADS_STEP / NL_BS)
nl_step_kernel(NlArgs a)
{
    __shared__ double scratch[NL_NS * (NL_BS / 32)];
    __shared__ bool is_last;
    __shared__ int s_viol;
    __shared__ int s_flag;
    (void)s_flag;
    constexpr int APB = NL_BS / TPA; // atoms per CTA
    const int n = a.sc.n;
    const int sub = threadIdx.x % TPA;
    int i = blockIdx.x * APB + threadIdx.x / TPA;
    bool valid = i < n;
    int nact = gridDim.x, seq = 0, sb = 0, se = n;
    if (SH) {
        seq = a.ctl->seq;
        // ONE thread decides for the whole CTA and broadcasts the decision through shared memory.
        // (Round 1 had every thread test ctl->halted itself: in the launch that raises the halt, CTA 0
        // sets the flag while other CTAs are starting, so warps of one CTA could disagree -- the warp of
        // thread 0 left, the others passed the barrier, read a never-written s_viol and advanced their
        // atoms by a step nobody else took.  That is the rare "wrong velocities in a block of CTAs
        // around a rebuild" of DESIGN.md section 6.)
        if (threadIdx.x < 32) { // warp 0 decides; lanes 0 .. world-1 wait for one rank's flag each, in parallel
            const int lane = threadIdx.x;
            int viol = 0;
            if (a.dbg_time && lane == 0 && seq < 1024) atomicMin(&a.dbg_time[4 * seq + 0], nl_globaltimer());
            if (lane == 0) {
                const int layer = a.nc * a.nc;
                const int owned = a.cell_start[a.xhi * layer] - a.cell_start[a.xlo * layer];
                if (a.ctl->halted || a.ctl->steps_left <= 0 || a.ctl->overflow != 0)
                    viol = 2; // nothing to do in this launch (a list overflow freezes the state at the last valid step)
                else if ((int)blockIdx.x >= (owned + APB - 1) / APB)
                    viol = 2; // CTA beyond the owned range (the grid is sized with slack): must not wait for flags
            }
            viol = __shfl_sync(0xffffffffu, viol, 0);
            if (viol == 0 && seq > 0) {
                // Wait for every rank's flag of step seq, but never forever: a peer that died, stopped
                // launching after an error or was given another step count would otherwise leave this GPU
                // spinning until a reset.  On expiry the run halts with NL_ERR_FLAG_TIMEOUT (the host
                // reports LJMD_ERR_CUDA); a failing rank also publishes a poison flag (NL_FLAG_POISON).
                int v = seq << 1;
                bool late = false;
                if (lane < a.world) {
                    const int *f = a.flags + (seq & 1) * LJMD_MAX_RANKS + lane;
                    const unsigned long long t0 = nl_globaltimer();
                    while (((v = nl_ld_acquire_sys(f)) >> 1) < seq) {
                        __nanosleep(40);
                        if (nl_globaltimer() - t0 > a.spin_budget_ns) { late = true; break; }
                    }
                }
                if (__any_sync(0xffffffffu, late)) { // timed out: freeze this rank (every CTA of this and later launches leaves at once)
                    if (lane == 0) { atomicMax(&a.ctl->overflow, NL_ERR_FLAG_TIMEOUT); a.ctl->halted = 1; }
                    viol = 2;
                } else if (__any_sync(0xffffffffu, (v & 1) != 0)) {
                    viol = 1;
                }
            }
            if (lane == 0) s_viol = viol;
            if (a.dbg_time && lane == 0 && viol == 0 && seq < 1024) atomicMax(&a.dbg_time[4 * seq + 1], nl_globaltimer());
        }
        __syncthreads();
        if (s_viol & 2) return;
        if (s_viol) { // some atom moved more than skin/2: the list may be stale for this step
            if (blockIdx.x == 0 && threadIdx.x == 0) { a.ctl->halted = 1; a.ctl->need_rebuild = 1; }
            return;
        }
        const int layer = a.nc * a.nc;
        sb = a.cell_start[a.xlo * layer];
        se = a.cell_start[a.xhi * layer];
        nact = (se - sb + APB - 1) / APB;
        if (nact > (int)gridDim.x) { // more owned atoms than the launch has CTAs for: flagged, reported by the host
            if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&a.ctl->overflow, NL_ERR_OWN_BLOCKS);
            nact = gridDim.x;
        }
        if ((int)blockIdx.x >= nact) return;
        i = sb + blockIdx.x * APB + threadIdx.x / TPA;
        valid = i < se;
        if (a.dbg_exec && threadIdx.x == 0 && seq < 4096) atomicAdd(&a.dbg_exec[seq], 1);
    }
    // A neighbour-list overflow (an atom with more neighbours than the capacity: its list was truncated)
    // freezes the state: the step is a no-op for every thread, so nothing is advanced with missing
    // interactions and the host finds the last valid step when it reports the error.  The flag is only
    // raised by the build kernels, which precede this launch in stream order: every thread reads the same value.
    if (!SH && MODE == 1 && a.ctl->overflow != 0) return;
    const int cur = a.ctl->cur;
    const int ev = (MODE == 1) ? 1 - cur : cur;
    const double4 *pos = a.p4[ev];
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;

    double xi = 0.0, yi = 0.0, zi = 0.0;
    int cnt = 0;
    bool inter = true;
    if (valid) {
        const double4 pi = nl_ld_pos(pos + i);
        xi = pi.x; yi = pi.y; zi = pi.z;
        cnt = a.nb_cnt[i];
        inter = a.interior[i] != 0;
    }
    const bool fast = __all_sync(0xffffffffu, inter) && (a.ctl->unwrapped == 0);
    double fx = 0.0, fy = 0.0, fz = 0.0, s12 = 0.0, s6 = 0.0;
    const int4 *nb = reinterpret_cast<const int4 *>(a.nbr) + i;
    if (PIPE) {
        if (fast) nl_walk_pipe<false>(nb, a.stride, cnt, pos, xi, yi, zi, L, invL, rc2, fx, fy, fz, s12, s6);
        else nl_walk_pipe<true>(nb, a.stride, cnt, pos, xi, yi, zi, L, invL, rc2, fx, fy, fz, s12, s6);
    } else {
        if (fast) nl_walk<false, TPA>(nb, a.stride, cnt, sub, pos, xi, yi, zi, L, invL, rc2, fx, fy, fz, s12, s6);
        else nl_walk<true, TPA>(nb, a.stride, cnt, sub, pos, xi, yi, zi, L, invL, rc2, fx, fy, fz, s12, s6);
    }
    if (TPA > 1) { // force partials of the TPA lanes of an atom, fixed butterfly order
#pragma unroll
        for (int o = 1; o < TPA; o <<= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
    }

    double s[NL_NS] = {s12, s6, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; // every lane contributes its own share of the pair sums
    double d2 = 0.0;
    bool pushed = false; // this thread stored a halo position into a peer rank's buffer
    (void)pushed;
    if (valid && sub == 0) {
        const size_t ix = i, iy = ix + n, iz = iy + n;
        const double axn = 24.0 * fx, ayn = 24.0 * fy, azn = 24.0 * fz; // lj_potential_energy.f90:189-191
        double vx = a.v[ix], vy = a.v[iy], vz = a.v[iz];
        if (MODE == 1) {
            const double axo = a.a[ix], ayo = a.a[iy], azo = a.a[iz];
            vx = vx + axo * a.sc.dt_half; vy = vy + ayo * a.sc.dt_half; vz = vz + azo * a.sc.dt_half; // verlet.f90:72-74
            vx = vx + axn * a.sc.dt_half; vy = vy + ayn * a.sc.dt_half; vz = vz + azn * a.sc.dt_half; // verlet.f90:86-88
            a.v[ix] = vx; a.v[iy] = vy; a.v[iz] = vz;
            double *pr = a.rb[cur]; // r(t): read for the unwrapped update, then overwritten with r(t+2dt)
            if (TRACK) { // md_simulation_program.f90:341-351 (dnint)
                double dx = xi - pr[ix], dy = yi - pr[iy], dz = zi - pr[iz];
                dx = dx - L * round(dx * invL);
                dy = dy - L * round(dy * invL);
                dz = dz - L * round(dz * invL);
                a.ru[ix] = a.ru[ix] + dx; a.ru[iy] = a.ru[iy] + dy; a.ru[iz] = a.ru[iz] + dz;
            }
            const double xn = nl_drift_wrap(xi, vx, axn, a.sc); // verlet.f90:58-63 of the next step
            const double yn = nl_drift_wrap(yi, vy, ayn, a.sc);
            const double zn = nl_drift_wrap(zi, vz, azn, a.sc);
            pr[ix] = xn; pr[iy] = yn; pr[iz] = zn;
            nl_st_pos(a.p4[cur] + i, xn, yn, zn);
            if (SH) {
                // halo push over NVLink: the first / last owned cell layer is the neighbour ranks' halo
                const int layer = a.nc * a.nc;
                const size_t off = (size_t)cur * n + i;
                if (i < a.cell_start[(a.xlo + 1) * layer]) {
                    const int left = (a.rank + a.world - 1) % a.world;
                    if (left != a.rank) { nl_st_pos(a.peer_p4[left] + off, xn, yn, zn); pushed = true; }
                }
                if (i >= a.cell_start[(a.xhi - 1) * layer]) {
                    const int right = (a.rank + 1) % a.world;
                    if (right != a.rank) { nl_st_pos(a.peer_p4[right] + off, xn, yn, zn); pushed = true; }
                }
            }
            d2 = nl_disp2(a, i, xn, yn, zn);
        }
        a.a[ix] = axn; a.a[iy] = ayn; a.a[iz] = azn;
        s[2] = vx * vx; s[3] = vy * vy; s[4] = vz * vz; // verlet.f90:93-95
        s[5] = vx; s[6] = vy; s[7] = vz;                // total momentum: an invariant (Newton 3), kept as a self-check
    }
    block_sum<NL_NS>(s, scratch);
    if (MODE == 1) nl_block_max(d2, scratch, a.ctl);
    if (SH) {
        // Every thread that stored into a peer's buffer makes those stores visible system-wide BEFORE it arrives at
        // the barrier below; thread 0 takes the ticket after the barrier, and the flag is published (release, system
        // scope) only after all tickets: a peer that acquires the flag sees every halo position.  Threads (and whole
        // CTAs: all but the two boundary layers) that pushed nothing skip the fence -- MEMBAR.SYS is not free.
        if (pushed) __threadfence_system();
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *sp = a.spart + (size_t)blockIdx.x * NL_NS;
#pragma unroll
        for (int q = 0; q < NL_NS; ++q) sp[q] = s[q];
        __threadfence();
        if (SH && a.dbg_time && seq < 1024) atomicMax(&a.dbg_time[4 * seq + 2], nl_globaltimer());
        const unsigned t = atomicAdd(&a.ctl->ticket_step, 1u);
        is_last = (t == (unsigned)nact - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // fixed-order reduction of all CTA partials
    // (the loads of four rows are issued before the first add: this loop is a chain of L2 round trips -- at
    // 8192 CTAs it was 25 us of a 294 us kernel -- the adds keep their order, so the sums are unchanged bit for bit)
    double r[NL_NS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int nrow = SH ? nact : (int)gridDim.x;
    // thread 0 needs two more values at the very end; their round trips overlap with the reduction
    long long step_pre = 0;
    unsigned long long maxd2_pre = 0ull;
    if (MODE == 1 && threadIdx.x == 0) {
        step_pre = *((volatile long long *)&a.sctl->step);
        maxd2_pre = *((volatile unsigned long long *)&a.ctl->maxd2_bits); // final: every CTA folded its maximum in before its ticket
    }
#if LJMD_NL_PUBLISH_IN_KERNEL
    if (SH && MODE == 1 && threadIdx.x < 32) {
        // Publish "step seq+1 done (+ skin violated)" to every rank NOW, before the scalar reduction below: all the
        // flag says is that every CTA of this rank has taken its ticket (so every halo position was fenced
        // system-wide before) and whether the skin was violated; the peers' next kernels then never wait for this
        // rank's bookkeeping.  One lane per peer: world release stores issued by one thread would queue up
        // behind each other's round trips.
        const unsigned long long mb = __shfl_sync(0xffffffffu, maxd2_pre, 0);
        const int flag_val = ((seq + 1) << 1) | (__longlong_as_double((long long)mb) > a.half_skin2 ? 1 : 0);
        if ((int)threadIdx.x < a.world)
            nl_st_release_sys(a.peer_flags[threadIdx.x] + ((seq + 1) & 1) * LJMD_MAX_RANKS + a.rank, flag_val);
    }
#endif
    for (int b = threadIdx.x; b < nrow; b += 4 * NL_BS) {
        double v[4][NL_NS];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = b + u * NL_BS;
#pragma unroll
            for (int q = 0; q < NL_NS; ++q) v[u][q] = row < nrow ? __ldcg(a.spart + (size_t)row * NL_NS + q) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (b + u * NL_BS < nrow) {
#pragma unroll
                for (int q = 0; q < NL_NS; ++q) r[q] += v[u][q];
            }
    }
    block_sum<NL_NS>(r, scratch);
    if (threadIdx.x == 0) {
        a.mom[0] = r[5]; a.mom[1] = r[6]; a.mom[2] = r[7]; // sum of v over the atoms this launch handled
        // ordered pairs: every unordered pair was counted twice
        const double S12 = 0.5 * r[0], S6 = 0.5 * r[1];
        double o[4];
        o[0] = 4.0 * (S12 - S6) + a.sc.tail_u;                    // :140,188,221
        o[1] = 0.5 * (r[2] + r[3] + r[4]);                        // verlet.f90:93-95
        o[2] = 24.0 * fma(-2.0, S12, S6) + a.sc.tail_du;          // :143,192,222
        o[3] = 24.0 * fma(26.0, S12, -7.0 * S6) + a.sc.tail_ddu;  // :178,193,223
        a.obs[0] = o[0]; a.obs[1] = o[1]; a.obs[2] = o[2]; a.obs[3] = o[3];
        if (MODE == 1) {
            StepCtl *sc = a.sctl;
            const long long step = step_pre;
            if (sc->series && step < sc->series_cap) {
                double *sr = sc->series + (size_t)step * 4;
                sr[0] = o[0]; sr[1] = o[1]; sr[2] = o[2]; sr[3] = o[3];
            }
            sc->step = step + 1;
            a.ctl->cur = 1 - cur;
            a.ctl->unwrapped = 0; // every position the next step reads has been wrapped
            const double m = __longlong_as_double((long long)maxd2_pre);
            if (!SH) {
                if (m > a.half_skin2) a.ctl->need_rebuild = 1;
            } else {
                // "step seq+1 done (+ skin violated)": published to every rank by nl_publish_kernel
                const int flag_val = ((seq + 1) << 1) | (m > a.half_skin2 ? 1 : 0);
                a.ctl->pending_flag = flag_val;

                a.ctl->seq = seq + 1;
                a.ctl->steps_left -= 1;
                a.ctl->steps_done += 1;
            }
        }
        a.ctl->ticket_step = 0;
        if (SH && a.dbg_time && seq < 1024) a.dbg_time[4 * seq + 3] = nl_globaltimer();
    }
}

// First step of a run: rb[1-cur] = wrap(drift(rb[cur], v, a)) (verlet.f90:58-63) + skin check.
__global__ void __launch_bounds__(NL_BS) nl_drift_kernel(NlArgs a)
{
    __shared__ double scratch[NL_BS / 32];
    const int n = a.sc.n;
    const int i = blockIdx.x * NL_BS + threadIdx.x;
    const int cur = a.ctl->cur;
    double d2 = 0.0;
    if (i < n) {
        const double *ps = a.rb[cur];
        double *pd = a.rb[1 - cur];
        const size_t ix = i, iy = ix + n, iz = iy + n;
        const double xn = nl_drift_wrap(ps[ix], a.v[ix], a.a[ix], a.sc);
        const double yn = nl_drift_wrap(ps[iy], a.v[iy], a.a[iy], a.sc);
        const double zn = nl_drift_wrap(ps[iz], a.v[iz], a.a[iz], a.sc);
        pd[ix] = xn; pd[iy] = yn; pd[iz] = zn;
        nl_st_pos(a.p4[1 - cur] + i, xn, yn, zn);
        d2 = nl_disp2(a, i, xn, yn, zn);
    }
    nl_block_max(d2, scratch, a.ctl);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&a.ctl->ticket_drift, 1u);
        if (t == gridDim.x - 1) {
            a.ctl->ticket_drift = 0;
            const double m = __longlong_as_double((long long)a.ctl->maxd2_bits);
            if (m > a.half_skin2) a.ctl->need_rebuild = 1;
        }
    }
}

// Sets the condition of a CUDA-graph IF node from the device flag (graph path only).
__global__ void nl_setcond_kernel(cudaGraphConditionalHandle h, const NlCtl *ctl)
{
    if (threadIdx.x == 0) cudaGraphSetConditional(h, ctl->need_rebuild ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------
// Order conversion (upload / download / tests)
// ------------------------------------------------------------------------------------------
// dst[q][orig[k]] = src_q[k]
__global__ void __launch_bounds__(256) nl_to_orig_kernel(int n, int narr, NlArgs a, double *dst)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int o = a.orig[k];
    for (int q = 0; q < narr; ++q) dst[(size_t)q * n + o] = a.arr[q][k];
}

// src = [r(3) | v(3) | a(3)][n] in ORIGINAL order -> the resident (cell-sorted) arrays, positions into rb[cur].
__global__ void __launch_bounds__(256) nl_from_orig_kernel(int n, NlArgs a, const double *src, int cur)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int o = a.orig[k];
    double *r = a.rb[cur];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[(size_t)c * n + k] = src[(size_t)c * n + o];
        a.v[(size_t)c * n + k] = src[(size_t)(3 + c) * n + o];
        a.a[(size_t)c * n + k] = src[(size_t)(6 + c) * n + o];
    }
}

__global__ void __launch_bounds__(256) nl_identity_kernel(int n, int *orig, int *slot_of_orig)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { orig[k] = k; slot_of_orig[k] = k; }
}

// cell id of every atom in ORIGINAL numbering, from the sorted tables
__global__ void __launch_bounds__(128) nl_cell_of_atom_kernel(int ncells, const int *start, const int *orig, int *out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    for (int k = start[c]; k < start[c + 1]; ++k) out[orig[k]] = c;
}

// in-cutoff / listed pair counts (ordered)
__global__ void __launch_bounds__(NL_BS) nl_count_kernel(NlArgs a)
{
    const int n = a.sc.n;
    const int i = blockIdx.x * NL_BS + threadIdx.x;
    unsigned long long ncand = 0, ncut = 0;
    if (i < n) {
        const double *px = a.rb[a.ctl->cur], *py = px + n, *pz = py + n;
        const double xi = px[i], yi = py[i], zi = pz[i];
        const int cnt = a.nb_cnt[i];
        for (int m = 0; m < cnt; ++m) {
            const int j = a.nbr[nl_entry(a.stride, i, m)];
            const double dx = mic(xi - px[j], a.sc.L, a.sc.invL);
            const double dy = mic(yi - py[j], a.sc.L, a.sc.invL);
            const double dz = mic(zi - pz[j], a.sc.L, a.sc.invL);
            ++ncand;
            if (dist2(dx, dy, dz) < a.sc.rc2) ++ncut;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
        ncut += __shfl_xor_sync(0xffffffffu, ncut, o);
    }
    if ((threadIdx.x & 31) == 0 && a.pair_counts) { atomicAdd(&a.pair_counts[0], ncand); atomicAdd(&a.pair_counts[1], ncut); }
}

// Debug: order-independent checksums (XOR of the bit patterns) of the complete state.
__global__ void __launch_bounds__(256) nl_checksum_kernel(NlArgs a, unsigned long long *out)
{
    const int n = a.sc.n;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long x[4] = {0ull, 0ull, 0ull, 0ull};
    if (k < n) {
        const int cur = a.ctl->cur;
        for (int c = 0; c < 3; ++c) {
            x[0] ^= (unsigned long long)__double_as_longlong(a.rb[1 - cur][(size_t)c * n + k]) * (k + 1);
            x[1] ^= (unsigned long long)__double_as_longlong(a.v[(size_t)c * n + k]) * (k + 1);
            x[2] ^= (unsigned long long)__double_as_longlong(a.a[(size_t)c * n + k]) * (k + 1);
        }
        const double4 p = a.p4[1 - cur][k];
        x[3] ^= ((unsigned long long)__double_as_longlong(p.x) ^ (unsigned long long)__double_as_longlong(p.y) * 3 ^ (unsigned long long)__double_as_longlong(p.z) * 5) * (k + 1);
    }
    for (int q = 0; q < 4; ++q) {
        for (int o = 16; o > 0; o >>= 1) x[q] ^= __shfl_xor_sync(0xffffffffu, x[q], o);
        if ((threadIdx.x & 31) == 0) atomicXor(out + q, x[q]);
    }
}

// Publishes the pending flag of the last completed step (what the first CTA of the next step
// kernel does); only needed when ranks are emulated by contexts of ONE process that are stepped
// alternately, where a peer's next launch cannot run while this one waits.
__global__ void nl_publish_kernel(NlArgs a)
{
    // one lane per peer: the system-scope release stores of a single thread would queue up behind each other
    // (each waits for the round trip of the one before), world of them in parallel cost one round trip
    if ((int)threadIdx.x >= a.world) return;
    const int seq = a.ctl->seq;
    if (seq <= 0 || a.ctl->halted) return;
    const int val = a.ctl->pending_flag;
    const int slot = (seq & 1) * LJMD_MAX_RANKS + a.rank;
    nl_st_release_sys(a.peer_flags[threadIdx.x] + slot, val);
}

// Debug: checksums keyed by the ORIGINAL atom index: invariant under a correct permutation.
__global__ void __launch_bounds__(256) nl_checksum_orig_kernel(NlArgs a, unsigned long long *out)
{
    const int n = a.sc.n;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long x[4] = {0ull, 0ull, 0ull, 0ull};
    if (k < n) {
        const unsigned long long key = (unsigned long long)a.orig[k] * 2654435761ull + 12345ull;
        for (int q = 0; q < 12; ++q) x[q / 3] ^= (unsigned long long)__double_as_longlong(a.arr[q][k]) * (key + q);
    }
    for (int q = 0; q < 4; ++q) {
        for (int o = 16; o > 0; o >>= 1) x[q] ^= __shfl_xor_sync(0xffffffffu, x[q], o);
        if ((threadIdx.x & 31) == 0) atomicXor(out + q, x[q]);
    }
}

// After a halt: forget the violation bits of the flags already received (both parity slots).
// Only the parity slot the halted step consumed is touched: the peers may already be writing the
// other one (a faster rank can finish its rebuild and its next step before this kernel runs).
__global__ void nl_clear_viol_kernel(int *flags, NlCtl *ctl)
{
    const int t = threadIdx.x;
    if (t < LJMD_MAX_RANKS) flags[(ctl->seq & 1) * LJMD_MAX_RANKS + t] &= ~1;
    if (t == 0) ctl->pending_flag &= ~1; // it is published again by the first step after the rebuild
}

// Slab sharding: owned slot ranges of the state arrays <-> a dense [narr][cmax] exchange buffer.
__global__ void __launch_bounds__(256) nl_pack_kernel(NlArgs a, double *buf, int cmax)
{
    const int layer = a.nc * a.nc;
    const int sb = a.cell_start[a.xlo * layer], cnt = a.cell_start[a.xhi * layer] - sb;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    for (int q = 0; q < a.narr; ++q) buf[(size_t)q * cmax + t] = a.arr[q][sb + t];
    buf[(size_t)a.narr * cmax + t] = (double)a.orig[sb + t]; // after slab-local rebuilds only the owner knows them
}

// recv = [world][narr + 1][cmax]; bounds[q] = first cell layer of rank q (bounds[world] = nc)
struct NlBounds { int x[LJMD_MAX_RANKS + 1]; };
__global__ void __launch_bounds__(256) nl_unpack_kernel(NlArgs a, const double *recv, int cmax, NlBounds bd)
{
    const int q = blockIdx.y;
    if (q == a.rank) return; // own range is already in place
    const int layer = a.nc * a.nc;
    const int sb = a.cell_start[bd.x[q] * layer], cnt = a.cell_start[bd.x[q + 1] * layer] - sb;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const double *src = recv + (size_t)q * (a.narr + 1) * cmax;
    for (int k = 0; k < a.narr; ++k) a.arr[k][sb + t] = src[(size_t)k * cmax + t];
    a.orig[sb + t] = (int)src[(size_t)a.narr * cmax + t];
}

// ==========================================================================================
// Host side
// ==========================================================================================
int nlist_cells_per_edge_host(const ljmd_params &p, double skin)
{
    // Largest nc with cell edge L/nc >= rc + skin (same back-off rule as the plain cell list).
    const double range = p.rc + skin;
    volatile double ratio = p.box_length / range;
    int nc = (int)floor(ratio);
    if (nc > 0 && (double)nc * range * (1.0 + 1e-12) > p.box_length) nc -= 1;
    if (nc > 128) nc = 128;
    return nc;
}

double nlist_default_skin()
{
    double skin = 0.3;
    if (const char *e = getenv("LJMD_NL_SKIN")) skin = atof(e);
    if (!(skin >= 0.0)) skin = 0.3;
    return skin;
}

static NlArgs make_args(ljmd_ctx *c, int eval_sel)
{
    const NlistWork &w = c->nw;
    NlArgs a;
    a.sc = c->sc;
    a.nc = w.nc; a.ncells = w.ncells; a.scale = w.scale;
    a.rl2 = w.rl2; a.half_skin2 = w.half_skin2;
    a.maxnb = w.maxnb; a.stride = w.stride;
    a.nzg = w.nzg; a.nunits = w.nunits; a.cap_stage = w.cap_stage;
    a.rb[0] = c->rbuf[0]; a.rb[1] = c->rbuf[1];
    a.p4[0] = w.p4; a.p4[1] = w.p4 + c->nR;
    a.v = c->st.vx; a.a = c->st.ax; a.ru = c->st.rux;
    a.r0 = w.r0; a.alt = w.alt;
    const size_t n = c->nR;
    int q = 0;
    for (int b = 0; b < 2; ++b) for (int k = 0; k < 3; ++k) a.arr[q++] = c->rbuf[b] + k * n;
    for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.vx + k * n;
    for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.ax + k * n;
    if (c->st.rux) for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.rux + k * n;
    a.narr = q;
    for (; q < 15; ++q) a.arr[q] = nullptr;
    a.cell_of_slot = w.cell_of_slot; a.cell_count = w.cell_count; a.cell_start = w.cell_start;
    a.cell_cursor = w.cell_cursor; a.tile_sum = w.tile_sum; a.keys = w.keys;
    a.orig = w.orig; a.orig_alt = w.orig_alt; a.slot_of_orig = w.slot_of_orig;
    a.interior = w.interior; a.nb_cnt = w.nb_cnt; a.nbr = w.nbr;
    a.ctl = w.ctl; a.sctl = c->fw.ctl;
    a.spart = w.spart; a.nblk = w.nblk;
    a.obs = c->fw.obs;
    a.mom = w.mom;
    a.eval_sel = eval_sel;
    a.pair_counts = nullptr;
    a.sh_on = w.sh_on; a.rank = w.rank; a.world = w.world; a.xlo = w.xlo; a.xhi = w.xhi;
    a.sl_on = 0; a.sl_rng = w.sl_rng;
    for (int q = 0; q < LJMD_MAX_RANKS; ++q) { a.peer_p4[q] = w.peer_p4[q]; a.peer_flags[q] = w.peer_flags[q]; }
    a.flags = w.flags;
    a.dbg_exec = w.dbg_exec;
    a.dbg_time = w.dbg_time;
    {
        const char *e = w.sh_on ? getenv("LJMD_SLAB_SPIN_TIMEOUT_S") : nullptr;
        const double budget_s = e ? atof(e) : 10.0;
        a.spin_budget_ns = (unsigned long long)((budget_s > 0.0 ? budget_s : 10.0) * 1e9);
    }
    if (w.sh_on) { a.sc.tail_u = 0.0; a.sc.tail_du = 0.0; a.sc.tail_ddu = 0.0; } // added once, after the cross-rank reduction
    return a;
}

int nlist_plan(ljmd_ctx *c)
{
    NlistWork &w = c->nw;
    if (c->R != 1) return set_error(LJMD_ERR_UNSUPPORTED, "neighbour list supports nreplicas == 1 only");
    const int n = c->sc.n;
    w.skin = nlist_default_skin();
    w.nc = nlist_cells_per_edge_host(c->p, w.skin);
    if (w.nc < 3)
        return set_error(LJMD_ERR_UNSUPPORTED, "neighbour list needs floor(L/(rc+skin)) >= 3 (got %d)", w.nc);
    w.ncells = w.nc * w.nc * w.nc;
    w.scale = (double)w.nc * c->p.inv_box_length;
    const double rl = c->p.rc + w.skin;
    w.rl2 = rl * rl;
    w.half_skin2 = 0.25 * w.skin * w.skin;
    const double expect = 4.18879020478639 * rl * rl * rl * c->p.density;
    int cap = (int)(1.6 * expect) + 24;
    if (const char *e = getenv("LJMD_NL_CAP")) cap = atoi(e);
    cap = std::max(8, std::min(cap, std::max(8, n - 1)));
    w.maxnb = (cap + 3) / 4 * 4;
    w.stride = ((size_t)n + 31) / 32 * 32;
    w.nblk = (n + NL_BS - 1) / NL_BS;
    // lanes per atom of the step kernel: 4 is opt-in (LJMD_NL_TPA=4): measured no gain at N=32k on one GPU
    // (30.0 vs 28.0 us); meant for the small slabs of 8-GPU runs, not yet measured there
    w.tpa = 1;
    w.pipe = (n <= 262144) ? 1 : 0; // software-pipelined gathers for small systems (few warps per SM)
    if (const char *e = getenv("LJMD_NL_PIPE")) w.pipe = atoi(e) != 0;
    if (const char *e = getenv("LJMD_NL_TPA")) w.tpa = (atoi(e) == 4) ? 4 : 1;
    w.narr = c->st.rux ? 15 : 12;
    // work units: runs of about NL_BS atoms along z inside one cell column
    const double apc = (double)n / w.ncells; // atoms per cell
    int G = (int)floor((double)NL_BS / apc);
    G = std::max(1, std::min(G, w.nc));
    w.nzg = (w.nc + G - 1) / G;
    w.nunits = w.nc * w.nc * w.nzg;
    const int gmax = (w.nc + w.nzg - 1) / w.nzg; // cells of the longest z-group
    int cap_stage = (int)(1.2 * 9.0 * std::min(w.nc, gmax + 2) * apc) + 96;
    if (const char *e = getenv("LJMD_NL_STAGE")) cap_stage = atoi(e);
    cap_stage = std::min(cap_stage, 9300); // 3 x 8 bytes per atom, 227 KB of shared memory per CTA
    cap_stage = std::min(cap_stage, 65535);
    w.cap_stage = (cap_stage + 15) / 16 * 16;
    w.smem_bytes = sizeof(double) * 3 * (size_t)w.cap_stage;
    LJMD_CUDA_CHECK(cudaFuncSetAttribute(nl_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem_bytes));

    w.cur_host = 0;
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_of_slot, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_count, sizeof(int) * (size_t)w.ncells));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_start, sizeof(int) * ((size_t)w.ncells + 1)));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_cursor, sizeof(int) * (size_t)w.ncells));
    LJMD_CUDA_CHECK(cudaMalloc(&w.tile_sum, sizeof(int) * (size_t)((w.ncells + 1023) / 1024)));
    LJMD_CUDA_CHECK(cudaMalloc(&w.keys, sizeof(unsigned long long) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.orig, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.orig_alt, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.slot_of_orig, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.interior, (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.nb_cnt, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.nbr, sizeof(int) * w.stride * (size_t)w.maxnb));
    LJMD_CUDA_CHECK(cudaMalloc(&w.p4, sizeof(double4) * 2 * (size_t)n));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.p4, 0, sizeof(double4) * 2 * (size_t)n, c->stream));
    LJMD_CUDA_CHECK(cudaMalloc(&w.r0, sizeof(double) * 3 * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.alt, sizeof(double) * 15 * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.spart, sizeof(double) * NL_NS * 4 * (size_t)w.nblk)); // room for 4 lanes per atom
    LJMD_CUDA_CHECK(cudaMalloc(&w.mom, sizeof(double) * 4));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.mom, 0, sizeof(double) * 4, c->stream));
    LJMD_CUDA_CHECK(cudaMalloc(&w.ctl, sizeof(NlCtl)));
    LJMD_CUDA_CHECK(cudaMalloc(&w.flags, sizeof(int) * 2 * LJMD_MAX_RANKS));
    LJMD_CUDA_CHECK(cudaMalloc(&w.sl_rng, sizeof(int) * 16));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.sl_rng, 0, sizeof(int) * 16, c->stream));
    if (getenv("LJMD_SLAB_TIMELINE")) {
        LJMD_CUDA_CHECK(cudaMalloc(&w.dbg_time, sizeof(unsigned long long) * 4096));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.dbg_time, 0, sizeof(unsigned long long) * 4096, c->stream));
    }
    if (getenv("LJMD_SLAB_DEBUG")) {
        LJMD_CUDA_CHECK(cudaMalloc(&w.dbg_exec, sizeof(int) * 4096));
        LJMD_CUDA_CHECK(cudaMemsetAsync(w.dbg_exec, 0, sizeof(int) * 4096, c->stream));
    }
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.flags, 0, sizeof(int) * 2 * LJMD_MAX_RANKS, c->stream));
    LJMD_CUDA_CHECK(cudaStreamCreateWithFlags(&w.side_stream, cudaStreamNonBlocking));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.cell_count, 0, sizeof(int) * (size_t)w.ncells, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.cell_cursor, 0, sizeof(int) * (size_t)w.ncells, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.nb_cnt, 0, sizeof(int) * (size_t)n, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.interior, 0, (size_t)n, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.r0, 0, sizeof(double) * 3 * (size_t)n, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.spart, 0, sizeof(double) * NL_NS * 4 * (size_t)w.nblk, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.ctl, 0, sizeof(NlCtl), c->stream));
    nl_identity_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, w.orig, w.slot_of_orig);
    LJMD_CUDA_CHECK(cudaGetLastError());
    LJMD_TRY(nlist_request_rebuild(c));
    return LJMD_OK;
}

void nlist_free(ljmd_ctx *c)
{
    NlistWork &w = c->nw;
    cudaFree(w.cell_of_slot); cudaFree(w.cell_count); cudaFree(w.cell_start); cudaFree(w.cell_cursor); cudaFree(w.tile_sum);
    cudaFree(w.keys); cudaFree(w.orig); cudaFree(w.orig_alt); cudaFree(w.slot_of_orig); cudaFree(w.interior);
    cudaFree(w.nb_cnt); cudaFree(w.nbr); cudaFree(w.p4); cudaFree(w.r0); cudaFree(w.alt); cudaFree(w.spart); cudaFree(w.mom); cudaFree(w.ctl);
    for (int q = 0; q < LJMD_MAX_RANKS; ++q) {
        if (w.ipc_p4[q]) cudaIpcCloseMemHandle(w.ipc_p4[q]);
        if (w.ipc_flags[q]) cudaIpcCloseMemHandle(w.ipc_flags[q]);
    }
    cudaFree(w.flags);
    cudaFree(w.sl_rng);
    if (w.side_stream) cudaStreamDestroy(w.side_stream);
    cudaFree(w.compat_r);
    if (w.ev_r_ready) cudaEventDestroy(w.ev_r_ready);
    if (w.ev_r_copied) cudaEventDestroy(w.ev_r_copied);
}

// Host-side write of selected NlCtl fields (stream-ordered).
static int ctl_poke(ljmd_ctx *c, size_t offset, const void *src, size_t bytes)
{
    // the source must outlive the async copy: stage through the pinned obs scratch
    char *stage = reinterpret_cast<char *>(c->h_obs) + sizeof(double) * 4 * (size_t)c->R;
    memcpy(stage, src, bytes);
    LJMD_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(c->nw.ctl) + offset, stage, bytes, cudaMemcpyHostToDevice,
                                    c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

int nlist_request_rebuild(ljmd_ctx *c)
{
    const int one = 1;
    return ctl_poke(c, offsetof(NlCtl, need_rebuild), &one, sizeof(int));
}

// The rebuild sequence (every kernel is a no-op unless the device flag is set).
int nlist_enqueue_rebuild(ljmd_ctx *c, int eval_sel)
{
    const NlistWork &w = c->nw;
    const int n = c->sc.n;
    const NlArgs a = make_args(c, eval_sel);
    nl_bin_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a);
    const int ntiles = (w.ncells + 1023) / 1024;
    nl_scan_tiles_kernel<<<ntiles, 1024, 0, c->stream>>>(a);
    nl_scan_sums_kernel<<<1, 1024, 0, c->stream>>>(a, ntiles);
    nl_scan_add_kernel<<<ntiles, 1024, 0, c->stream>>>(a);
    nl_fill_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a);
    nl_sortcell_kernel<<<(w.ncells + 127) / 128, 128, 0, c->stream>>>(a);
    nl_permute_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a);
    nl_copyback_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a);
    nl_build_kernel<<<w.nunits, NL_BS, w.smem_bytes, c->stream>>>(a);
    if (!c->capturing) c->launches += 9;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

int nlist_enqueue_forces(ljmd_ctx *c)
{
    LJMD_TRY(nlist_enqueue_rebuild(c, 0));
    const NlArgs a = make_args(c, 0);
    if (c->nw.tpa == 4) nl_step_kernel<0, false, false, 4><<<(c->sc.n + NL_BS / 4 - 1) / (NL_BS / 4), NL_BS, 0, c->stream>>>(a);
    else nl_step_kernel<0, false, false, 1><<<c->nw.nblk, NL_BS, 0, c->stream>>>(a);
    if (!c->capturing) c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

int nlist_enqueue_drift(ljmd_ctx *c)
{
    const NlArgs a = make_args(c, 0);
    nl_drift_kernel<<<c->nw.nblk, NL_BS, 0, c->stream>>>(a);
    if (!c->capturing) c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

int nlist_enqueue_step_kernel(ljmd_ctx *c)
{
    const NlArgs a = make_args(c, 1);
    const bool track = c->st.rux != nullptr;
    const int tpa = c->nw.tpa;
    const int grid = c->nw.sh_on ? c->nw.own_blocks : (c->sc.n + NL_BS / tpa - 1) / (NL_BS / tpa);
#define LJMD_NL_LAUNCH(TR, SHD, T) nl_step_kernel<1, TR, SHD, T><<<grid, NL_BS, 0, c->stream>>>(a)
#define LJMD_NL_LAUNCH_PIPE(TR, SHD) nl_step_kernel<1, TR, SHD, 1, true><<<grid, NL_BS, 0, c->stream>>>(a)
    if (c->nw.pipe && tpa == 1) { // small systems / small slabs: pipelined gathers
        if (c->nw.sh_on) { if (track) LJMD_NL_LAUNCH_PIPE(true, true); else LJMD_NL_LAUNCH_PIPE(false, true); }
        else { if (track) LJMD_NL_LAUNCH_PIPE(true, false); else LJMD_NL_LAUNCH_PIPE(false, false); }
    } else if (c->nw.sh_on) {
        if (tpa == 4) { if (track) LJMD_NL_LAUNCH(true, true, 4); else LJMD_NL_LAUNCH(false, true, 4); }
        else { if (track) LJMD_NL_LAUNCH(true, true, 1); else LJMD_NL_LAUNCH(false, true, 1); }
    } else {
        if (tpa == 4) { if (track) LJMD_NL_LAUNCH(true, false, 4); else LJMD_NL_LAUNCH(false, false, 4); }
        else { if (track) LJMD_NL_LAUNCH(true, false, 1); else LJMD_NL_LAUNCH(false, false, 1); }
    }
#undef LJMD_NL_LAUNCH
#undef LJMD_NL_LAUNCH_PIPE
    if (!c->capturing) c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    c->nw.cur_host ^= 1;
    return LJMD_OK;
}

// One velocity-Verlet step (rb[1-cur] must already hold the drifted positions).
int nlist_enqueue_step(ljmd_ctx *c)
{
    LJMD_TRY(nlist_enqueue_rebuild(c, 1));
    return nlist_enqueue_step_kernel(c);
}

// `nsteps` steps as a CUDA graph in which the rebuild sequence of every step sits inside an IF
// node whose condition is the device flag: a step that needs no rebuild costs two launches
// (the 1-thread condition kernel and the step kernel).
int nlist_capture_graph(ljmd_ctx *c, int nsteps, cudaGraphExec_t *out)
{
    cudaGraph_t g = nullptr;
    LJMD_CUDA_CHECK(cudaGraphCreate(&g, 0));
    cudaStream_t s = c->stream;
    cudaError_t e = cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGraphDestroy(g); LJMD_CUDA_CHECK(e); }
    c->capturing = true;
    int rc = LJMD_OK;
    const int cur_save = c->nw.cur_host;
    std::vector<cudaGraph_t> bodies;
    for (int st = 0; st < nsteps && rc == LJMD_OK; ++st) {
        cudaGraphConditionalHandle h;
        e = cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
        if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "cudaGraphConditionalHandleCreate: %s", cudaGetErrorString(e)); break; }
        nl_setcond_kernel<<<1, 32, 0, s>>>(h, c->nw.ctl);
        cudaStreamCaptureStatus status;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        e = cudaStreamGetCaptureInfo(s, &status, nullptr, nullptr, &deps, &ndeps);
        if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "cudaStreamGetCaptureInfo: %s", cudaGetErrorString(e)); break; }
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = h;
        cp.conditional.type = cudaGraphCondTypeIf;
        cp.conditional.size = 1;
        cudaGraphNode_t cnode;
        e = cudaGraphAddNode(&cnode, g, deps, ndeps, &cp);
        if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "cudaGraphAddNode(conditional): %s", cudaGetErrorString(e)); break; }
        bodies.push_back(cp.conditional.phGraph_out[0]);
        e = cudaStreamUpdateCaptureDependencies(s, &cnode, 1, cudaStreamSetCaptureDependencies);
        if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "cudaStreamUpdateCaptureDependencies: %s", cudaGetErrorString(e)); break; }
        rc = nlist_enqueue_step_kernel(c);
    }
    cudaGraph_t g2 = nullptr;
    e = cudaStreamEndCapture(s, &g2);
    if (rc == LJMD_OK && e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    // bodies of the IF nodes: the rebuild sequence, captured on a side stream
    for (size_t b = 0; b < bodies.size() && rc == LJMD_OK; ++b) {
        e = cudaStreamBeginCaptureToGraph(c->nw.side_stream, bodies[b], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "capture of the rebuild body: %s", cudaGetErrorString(e)); break; }
        c->stream = c->nw.side_stream;
        rc = nlist_enqueue_rebuild(c, 1);
        c->stream = s;
        cudaGraph_t dummy = nullptr;
        e = cudaStreamEndCapture(c->nw.side_stream, &dummy);
        if (rc == LJMD_OK && e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "end capture of the rebuild body: %s", cudaGetErrorString(e));
    }
    c->capturing = false;
    c->nw.cur_host = cur_save; // capture does not execute anything
    if (rc != LJMD_OK) { cudaGraphDestroy(g); return rc; }
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    LJMD_CUDA_CHECK(e);
    return LJMD_OK;
}

// State in ORIGINAL order into w.alt: arrays [r(3) | v(3) | a(3) | ru(3)] (positions = rb[cur]).
int nlist_gather_orig(ljmd_ctx *c)
{
    const NlistWork &w = c->nw;
    const int n = c->sc.n;
    NlArgs a = make_args(c, 0);
    const size_t nn = c->nR;
    int q = 0;
    for (int k = 0; k < 3; ++k) a.arr[q++] = c->rbuf[w.cur_host] + k * nn;
    for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.vx + k * nn;
    for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.ax + k * nn;
    if (c->st.rux) for (int k = 0; k < 3; ++k) a.arr[q++] = c->st.rux + k * nn;
    nl_to_orig_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, q, a, w.alt);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// Drop-in calls with page-locked caller arrays (large systems; the call is PCIe bound: 75 MB each way at N = 2^20
// against 0.3 ms of kernels).  r(t+dt) is final as soon as the drift has run -- BEFORE the force / kick kernel --
// so it is gathered into its own buffer (original order; w.alt is the permutation scratch of a list rebuild that
// the step may trigger, so it cannot be used) and its device-to-host copies run on the side stream while the
// main stream computes the step; v and a follow on the main stream afterwards.  Called between the drift and the
// step: rb[1 - cur] holds r(t+dt) and `orig` is the table of the order it is stored in (a rebuild inside the step
// changes both, after this gather in stream order).  The caller makes the main stream wait for ev_r_copied.
// Measured at N = 2^20 (profiles/r2_e2e_overlap.txt): 3.36 -> 3.15 ms per call on one box, 2.89 ms on another.  Going
// further -- upload, drift and this download pipelined over 4 / 8 chunks of atoms, PCIe in both directions at once --
// was built, bitwise equal, and SLOWER (3.16 / 3.36 ms against 2.89 ms on the same box): 12 more copies per chunk at
// ~7 us of fixed cost each outweigh the hidden 0.46 ms.  Removed again.
int nlist_early_positions_d2h(ljmd_ctx *c, double *const host_r[3])
{
    NlistWork &w = c->nw;
    const int n = c->sc.n;
    const size_t nn = c->nR, bytes = sizeof(double) * nn;
    if (!w.compat_r) LJMD_CUDA_CHECK(cudaMalloc(&w.compat_r, 3 * bytes));
    if (!w.ev_r_ready) LJMD_CUDA_CHECK(cudaEventCreateWithFlags(&w.ev_r_ready, cudaEventDisableTiming));
    if (!w.ev_r_copied) LJMD_CUDA_CHECK(cudaEventCreateWithFlags(&w.ev_r_copied, cudaEventDisableTiming));
    NlArgs a = make_args(c, 0);
    for (int k = 0; k < 3; ++k) a.arr[k] = c->rbuf[1 - w.cur_host] + k * nn;
    nl_to_orig_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, 3, a, w.compat_r);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    LJMD_CUDA_CHECK(cudaEventRecord(w.ev_r_ready, c->stream));
    LJMD_CUDA_CHECK(cudaStreamWaitEvent(w.side_stream, w.ev_r_ready, 0));
    for (int k = 0; k < 3; ++k)
        LJMD_CUDA_CHECK(cudaMemcpyAsync(host_r[k], w.compat_r + (size_t)k * nn, bytes, cudaMemcpyDeviceToHost, w.side_stream));
    LJMD_CUDA_CHECK(cudaEventRecord(w.ev_r_copied, w.side_stream));
    return LJMD_OK;
}

// v and a in ORIGINAL order into w.alt[3..8] (the layout of nlist_gather_orig without the positions).
int nlist_gather_orig_va(ljmd_ctx *c)
{
    const int n = c->sc.n;
    const size_t nn = c->nR;
    NlArgs a = make_args(c, 0);
    for (int k = 0; k < 3; ++k) a.arr[k] = c->st.vx + k * nn;
    for (int k = 0; k < 3; ++k) a.arr[3 + k] = c->st.ax + k * nn;
    nl_to_orig_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, 6, a, c->nw.alt + 3 * nn);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// The drop-in call keeps the cell-sorted order and the neighbour list of the previous call: the caller's arrays
// (original order, already copied to w.alt as [r | v | a]) are scattered into the resident order.  Whether the
// list is still good for them is decided by the usual skin test of the drift that follows (displacement of
// r(t+dt) from the build positions): consecutive calls of an MD loop reuse the list ~40 times, anything else
// (another configuration, permuted atoms) moves some atom by more than skin/2 and triggers a rebuild.
int nlist_load_from_orig(ljmd_ctx *c)
{
    const int n = c->sc.n;
    const NlArgs a = make_args(c, 0);
    nl_from_orig_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, a, c->nw.alt, c->nw.cur_host);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// Puts the resident state back into ORIGINAL order with cur = 0 (before an upload).
int nlist_canonicalize(ljmd_ctx *c)
{
    NlistWork &w = c->nw;
    const size_t nn = c->nR, bytes = sizeof(double) * nn;
    LJMD_TRY(nlist_gather_orig(c));
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->rbuf[0], w.alt, 3 * bytes, cudaMemcpyDeviceToDevice, c->stream));
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.vx, w.alt + 3 * nn, 3 * bytes, cudaMemcpyDeviceToDevice, c->stream));
    LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.ax, w.alt + 6 * nn, 3 * bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (c->st.rux) LJMD_CUDA_CHECK(cudaMemcpyAsync(c->st.rux, w.alt + 9 * nn, 3 * bytes, cudaMemcpyDeviceToDevice, c->stream));
    return nlist_reset_identity(c); // keeps the rebuild counter
}

// The state arrays were just overwritten in ORIGINAL order (drop-in calls): identity permutation,
// cur = 0, list invalid.
int nlist_reset_identity(ljmd_ctx *c)
{
    NlistWork &w = c->nw;
    const int n = c->sc.n;
    nl_identity_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, w.orig, w.slot_of_orig);
    LJMD_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    NlCtl h;
    memset(&h, 0, sizeof(h));
    h.need_rebuild = 1;
    LJMD_TRY(ctl_poke(c, 0, &h, offsetof(NlCtl, rebuilds)));
    w.cur_host = 0;
    return LJMD_OK;
}

int nlist_check_overflow(ljmd_ctx *c)
{
    NlCtl h;
    LJMD_CUDA_CHECK(cudaMemcpyAsync(&h, c->nw.ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->nw.rebuilds_host = h.rebuilds;
    if (h.overflow != 0) c->nw.cur_host = h.cur; // step launches after the error were no-ops: follow the device
    if (h.overflow >= NL_ERR_OWN_BLOCKS)
        return set_error(LJMD_ERR_UNSUPPORTED, "slab sharding: a rank owns more atoms than its launch capacity (density too inhomogeneous)");
    if (h.overflow >= NL_ERR_FLAG_TIMEOUT)
        return set_error(LJMD_ERR_CUDA, "slab sharding: a peer rank's per-step flag did not arrive within the spin budget "
                                        "(LJMD_SLAB_SPIN_TIMEOUT_S); the run is halted, the state is at the last completed step");
    if (h.overflow > 0)
        return set_error(LJMD_ERR_UNSUPPORTED,
                         "neighbour list capacity exceeded: an atom has %d neighbours within rc+skin, capacity %d "
                         "(set LJMD_NL_CAP); stepping stopped at the last valid step -- re-create the context with a "
                         "larger capacity and upload again", h.overflow, c->nw.maxnb);
    return LJMD_OK;
}

int nlist_pair_counts(ljmd_ctx *c, unsigned long long *d_counts)
{
    LJMD_TRY(nlist_enqueue_rebuild(c, 0));
    NlArgs a = make_args(c, 0);
    a.pair_counts = d_counts;
    nl_count_kernel<<<c->nw.nblk, NL_BS, 0, c->stream>>>(a);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

int nlist_download_cells(ljmd_ctx *c, int32_t *cell_of_atom, int32_t *cell_start, int32_t *order)
{
    const NlistWork &w = c->nw;
    const int n = c->sc.n;
    cudaError_t e = cudaSuccess;
    if (cell_of_atom) {
        // cell_of_slot is scratch between rebuilds: reuse it for the original-numbering table
        nl_cell_of_atom_kernel<<<(w.ncells + 127) / 128, 128, 0, c->stream>>>(w.ncells, w.cell_start, w.orig, w.cell_of_slot);
        e = cudaMemcpyAsync(cell_of_atom, w.cell_of_slot, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess && cell_start)
        e = cudaMemcpyAsync(cell_start, w.cell_start, sizeof(int) * ((size_t)w.ncells + 1), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && order)
        e = cudaMemcpyAsync(order, w.orig, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_download_cells(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// ------------------------------------------------------------------------------------------
// Slab sharding of one large system over the GPUs of a node (one process per GPU).
// ------------------------------------------------------------------------------------------
static void slab_layers(int nc, int rank, int world, int *xlo, int *xhi)
{
    *xlo = (int)((long long)rank * nc / world);
    *xhi = (int)((long long)(rank + 1) * nc / world);
}

int nlist_slab_bounds_host(ljmd_ctx *c, int32_t *bounds)
{
    const NlistWork &w = c->nw;
    const int layer = w.nc * w.nc;
    for (int q = 0; q <= w.world; ++q) {
        int xlo, xhi;
        slab_layers(w.nc, q < w.world ? q : 0, w.world, &xlo, &xhi);
        const int cell = (q < w.world) ? xlo * layer : w.ncells;
        LJMD_CUDA_CHECK(cudaMemcpyAsync(&bounds[q], w.cell_start + cell, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    LJMD_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return LJMD_OK;
}

} // namespace ljmd

using namespace ljmd;

extern "C" int ljmd_ctx_slab_export(ljmd_ctx *c, void *handles128, uint64_t *raw_ptrs2)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_export(): ctx is NULL.");
    if (c->method != LJMD_METHOD_NLIST) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_slab_export(): needs the neighbour-list method.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    cudaError_t e = cudaSuccess;
    if (handles128) {
        cudaIpcMemHandle_t h[2];
        e = cudaIpcGetMemHandle(&h[0], c->nw.p4);
        if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[1], c->nw.flags);
        if (e == cudaSuccess) memcpy(handles128, h, sizeof(h));
    }
    if (raw_ptrs2) { raw_ptrs2[0] = (uint64_t)(uintptr_t)c->nw.p4; raw_ptrs2[1] = (uint64_t)(uintptr_t)c->nw.flags; }
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_export(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_slab_attach(ljmd_ctx *c, int32_t rank, int32_t world, const void *all_handles,
                                    const uint64_t *all_raw_ptrs)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_attach(): ctx is NULL.");
    if (c->method != LJMD_METHOD_NLIST) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_slab_attach(): needs the neighbour-list method.");
    if (world < 1 || world > LJMD_MAX_RANKS || rank < 0 || rank >= world)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_attach(): bad rank %d / world %d (max %d).", rank, world, LJMD_MAX_RANKS);
    NlistWork &w = c->nw;
    if (w.nc < world) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_slab_attach(): %d cell layers cannot be split over %d ranks.", w.nc, world);
    if (!all_handles && !all_raw_ptrs) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_attach(): no peer handles.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    int rc = LJMD_OK;
    for (int q = 0; q < world && rc == LJMD_OK; ++q) {
        if (q == rank) { w.peer_p4[q] = w.p4; w.peer_flags[q] = w.flags; continue; }
        if (all_raw_ptrs) { // same process (tests): the peers' device pointers are directly usable
            w.peer_p4[q] = reinterpret_cast<double4 *>((uintptr_t)all_raw_ptrs[2 * q]);
            w.peer_flags[q] = reinterpret_cast<int *>((uintptr_t)all_raw_ptrs[2 * q + 1]);
        } else {
            const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles) + 2 * q;
            void *p = nullptr, *f = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h[0], cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&f, h[1], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { rc = set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_attach(): cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e)); break; }
            w.ipc_p4[q] = p; w.ipc_flags[q] = f;
            w.peer_p4[q] = reinterpret_cast<double4 *>(p);
            w.peer_flags[q] = reinterpret_cast<int *>(f);
        }
    }
    if (rc == LJMD_OK) {
        w.sh_on = 1; w.rank = rank; w.world = world;
        slab_layers(w.nc, rank, world, &w.xlo, &w.xhi);
        // CTAs of the step kernel: room for 30 % more atoms than the mean slab (+ one layer's worth)
        const double mean = (double)c->sc.n * (w.xhi - w.xlo) / w.nc;
        // sharded runs: the plain walk at 8 CTAs per SM (the pipelined variant holds 4 CTAs per SM: two waves for an
        // 8-GPU slab of N = 2^20, and it measured 14 % slower at 2 GPUs, profiles/r2_slab_variants_n2.txt)
        w.pipe = 0;
        if (const char *e = getenv("LJMD_NL_PIPE")) w.pipe = atoi(e) != 0;
        const int apb = NL_BS / w.tpa;
        w.own_blocks = (int)((1.3 * mean + 4096.0) / apb) + 1;
        if (w.own_blocks > w.nblk * w.tpa) w.own_blocks = w.nblk * w.tpa;
        // slab-local rebuild: two cell layers of a neighbour (30 % above the mean + slack), and all candidates
        w.sl_halo_cap = (int)(2.0 * (1.3 * (double)c->sc.n / w.nc + 1024.0));
        w.sl_cand_cap = std::min(c->sc.n, w.own_blocks * apb + 2 * w.sl_halo_cap);
        if (c->graph_chunk) { cudaGraphExecDestroy(c->graph_chunk); c->graph_chunk = nullptr; }
        for (int k = 0; k < 33; ++k)
            if (c->graph_slab[k]) { cudaGraphExecDestroy(c->graph_slab[k]); c->graph_slab[k] = nullptr; }
        int zero4[5] = {0, 0, 0, 0, 0}; // halted, steps_left, steps_done, seq, pending_flag
        cudaError_t e = cudaMemcpyAsync(reinterpret_cast<char *>(w.ctl) + offsetof(NlCtl, halted), zero4, sizeof(zero4),
                                        cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(w.flags, 0, sizeof(int) * 2 * LJMD_MAX_RANKS, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_attach(): %s", cudaGetErrorString(e));
    }
    if (prev != c->device) cudaSetDevice(prev);
    return rc;
}

// First slot of every rank's slab (bounds[world] = n), from the current cell table.
extern "C" int ljmd_ctx_slab_bounds(ljmd_ctx *c, int32_t *bounds)
{
    clear_error();
    if (!c || !bounds) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_bounds(): NULL argument.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_bounds(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const int rc = nlist_slab_bounds_host(c, bounds);
    if (prev != c->device) cudaSetDevice(prev);
    return rc;
}

// Owned slot range of every permuted state array -> dense device buffer [narr][cmax].
extern "C" int ljmd_ctx_slab_pack(ljmd_ctx *c, double *dev_buf, int32_t cmax, int32_t *narr)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_pack(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_pack(): context is not sharded.");
    if (narr) *narr = c->nw.narr + 1; // the state arrays + the original atom indices
    if (!dev_buf) return LJMD_OK; // query only
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const NlArgs a = make_args(c, 0);
    nl_pack_kernel<<<(cmax + 255) / 256, 256, 0, c->stream>>>(a, dev_buf, cmax);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_pack(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// All ranks' packed ranges ([world][narr][cmax]) -> the state arrays (every rank then holds the
// complete state).
extern "C" int ljmd_ctx_slab_unpack(ljmd_ctx *c, const double *dev_buf, int32_t cmax)
{
    clear_error();
    if (!c || !dev_buf) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_unpack(): NULL argument.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_unpack(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const NlArgs a = make_args(c, 0);
    NlBounds bd;
    for (int q = 0; q <= c->nw.world; ++q) {
        int xlo, xhi;
        slab_layers(c->nw.nc, q < c->nw.world ? q : 0, c->nw.world, &xlo, &xhi);
        bd.x[q] = (q < c->nw.world) ? xlo : c->nw.nc;
    }
    dim3 grid((cmax + 255) / 256, c->nw.world);
    nl_unpack_kernel<<<grid, 256, 0, c->stream>>>(a, dev_buf, cmax, bd);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_unpack(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// Global re-sort + list rebuild for the positions of the NEXT step if the device flag asks for it
// (a halt sets it) or `force` is non-zero; clears the halt.  Every rank must hold the complete
// state: call after ljmd_ctx_slab_unpack.
extern "C" int ljmd_ctx_slab_rebuild(ljmd_ctx *c, int32_t force)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_rebuild(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_rebuild(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    nl_clear_viol_kernel<<<1, 32, 0, c->stream>>>(c->nw.flags, c->nw.ctl);
    int rc = force ? nlist_request_rebuild(c) : LJMD_OK;
    const int zero = 0;
    if (rc == LJMD_OK) {
        cudaError_t e = cudaMemcpyAsync(reinterpret_cast<char *>(c->nw.ctl) + offsetof(NlCtl, halted), &zero, sizeof(int),
                                        cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_rebuild(): %s", cudaGetErrorString(e));
    }
    unsigned long long *dsum = nullptr;
    if (c->nw.dbg_exec && getenv("LJMD_SLAB_DEBUG_SUM")) { // debug: the permutation must preserve every per-atom value
        cudaMalloc(&dsum, 64);
        cudaMemsetAsync(dsum, 0, 64, c->stream);
        const NlArgs a = make_args(c, 1);
        nl_checksum_orig_kernel<<<(c->sc.n + 255) / 256, 256, 0, c->stream>>>(a, dsum);
    }
    if (rc == LJMD_OK) rc = nlist_enqueue_rebuild(c, 1);
    if (dsum) {
        const NlArgs a = make_args(c, 1);
        nl_checksum_orig_kernel<<<(c->sc.n + 255) / 256, 256, 0, c->stream>>>(a, dsum + 4);
        unsigned long long hs[8];
        cudaMemcpyAsync(hs, dsum, 64, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(dsum);
        const char *names[4] = {"rb0", "rb1", "v", "a"};
        for (int q = 0; q < 4; ++q)
            if (hs[q] != hs[4 + q])
                fprintf(stderr, "[ljmd slab debug] rank %d: rebuild changed the per-atom checksum of %s\n", c->nw.rank, names[q]);
    }
    if (rc == LJMD_OK) rc = nlist_check_overflow(c);
    if (prev != c->device) cudaSetDevice(prev);
    return rc;
}

// ------------------------------------------------------------------------------------------
// Slab-LOCAL rebuild: no rank touches more than its own slab and two cell layers of each neighbour.
//   1. ljmd_ctx_slab_halo(pack, side 0 / 1): my first / last two cell layers (all permuted state arrays + original
//      indices) -> send buffers; the host exchanges them with the two neighbours (NCCL send / recv);
//      ljmd_ctx_slab_halo(unpack, side 0 / 1): the left neighbour's last / the right neighbour's first two layers
//      -> their global slots in this rank's arrays.  The cell table is still the old, globally identical one, so
//      both sides agree on every count and slot.
//   2. ljmd_ctx_slab_rebuild_local(0): cell of every CANDIDATE (own slab + the four received layers: every atom
//      that can now be in an owned cell or in the two halo layers, because nobody moved more than skin/2 < one cell);
//      histogram of the OWN layers' cells.  Host: ONE all-reduce (sum, int32) of the histogram over the ranks
//      (ljmd_ctx_slab_local_info gives the device address) -> the complete, globally identical new histogram.
//   3. ljmd_ctx_slab_rebuild_local(1): exclusive scan (-> new global cell table), scatter + per-cell sort by original
//      index (the stable counting sort: same slots as the global sort, bit-exact) of the own and halo cells,
//      permutation of their state, neighbour lists of the own slab.
// Afterwards this rank's arrays are current for the own slab and the two halo layers only; ljmd_ctx_slab_pack /
// _unpack (a full gather) make them complete again (SlabRunner.finish()).  Needs >= 2 cell layers per rank.
// ------------------------------------------------------------------------------------------
extern "C" int ljmd_ctx_slab_local_info(ljmd_ctx *c, uint64_t *cell_count_ptr, int32_t *ncells, int32_t *halo_cap,
                                        int32_t *halo_narr)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_local_info(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_local_info(): context is not sharded.");
    if (c->nw.world < 2 || c->nw.nc < 2 * c->nw.world)
        return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_slab_local_info(): a slab-local rebuild needs >= 2 ranks and >= 2 cell layers per rank.");
    if (cell_count_ptr) *cell_count_ptr = (uint64_t)(uintptr_t)c->nw.cell_count;
    if (ncells) *ncells = c->nw.ncells;
    if (halo_cap) *halo_cap = c->nw.sl_halo_cap;
    if (halo_narr) *halo_narr = c->nw.narr + 1;
    return LJMD_OK;
}

extern "C" int ljmd_ctx_slab_halo(ljmd_ctx *c, double *dev_buf, int32_t cap, int32_t side, int32_t unpack)
{
    clear_error();
    if (!c || !dev_buf) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_halo(): NULL argument.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_halo(): context is not sharded.");
    if (cap < c->nw.sl_halo_cap) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_halo(): cap %d < %d.", cap, c->nw.sl_halo_cap);
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const NlistWork &w = c->nw;
    const NlArgs a = make_args(c, 1);
    int l0;
    if (!unpack) l0 = side == 0 ? w.xlo : w.xhi - 2;                                  // mine: first / last two layers
    else l0 = side == 0 ? ((w.xlo - 2) % w.nc + w.nc) % w.nc : w.xhi % w.nc;          // the neighbours'
    nl_halo_pack_kernel<<<(cap + 255) / 256, 256, 0, c->stream>>>(a, dev_buf, cap, l0, unpack ? 1 : 0);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_halo(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_slab_rebuild_local(ljmd_ctx *c, int32_t phase)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_rebuild_local(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_rebuild_local(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    NlistWork &w = c->nw;
    NlArgs a = make_args(c, 1);
    a.sl_on = 1;
    int rc = LJMD_OK;
    cudaError_t e = cudaSuccess;
    const int nb_cand = (w.sl_cand_cap + 255) / 256;
    if (phase == 0) {
        nl_clear_viol_kernel<<<1, 32, 0, c->stream>>>(w.flags, w.ctl);
        const int zero = 0; // the halt is over; async copy from pageable memory returns after staging
        e = cudaMemcpyAsync(reinterpret_cast<char *>(w.ctl) + offsetof(NlCtl, halted), &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream);
        nl_sl_ranges_kernel<<<1, 32, 0, c->stream>>>(a, w.sl_rng, 0);
        nl_bin_kernel<<<nb_cand, 256, 0, c->stream>>>(a);
        c->launches += 3;
    } else {
        const int ntiles = (w.ncells + 1023) / 1024;
        const int ncell_tgt = (w.xhi - w.xlo + 2) * w.nc * w.nc;
        nl_scan_tiles_kernel<<<ntiles, 1024, 0, c->stream>>>(a);
        nl_scan_sums_kernel<<<1, 1024, 0, c->stream>>>(a, ntiles);
        nl_scan_add_kernel<<<ntiles, 1024, 0, c->stream>>>(a);
        nl_sl_ranges_kernel<<<1, 32, 0, c->stream>>>(a, w.sl_rng, 1);
        nl_fill_kernel<<<nb_cand, 256, 0, c->stream>>>(a);
        nl_sortcell_kernel<<<(ncell_tgt + 127) / 128, 128, 0, c->stream>>>(a);
        nl_permute_kernel<<<nb_cand, 256, 0, c->stream>>>(a);
        nl_copyback_kernel<<<nb_cand, 256, 0, c->stream>>>(a);
        nl_build_kernel<<<w.nunits, NL_BS, w.smem_bytes, c->stream>>>(a);
        c->launches += 9;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_rebuild_local(): %s", cudaGetErrorString(e));
    // no host synchronisation here: a list overflow of this build is reported by the next ljmd_ctx_slab_run
    // (its step kernels are no-ops once the flag is set), so the host can queue the next chunk at once
    if (prev != c->device) cudaSetDevice(prev);
    return rc;
}

// Captures the CUDA graph of a chunk of max_steps (2 .. 32) sharded steps now, so that the first
// ljmd_ctx_slab_run of that length does not pay for it.
extern "C" int ljmd_ctx_slab_prepare(ljmd_ctx *c, int32_t max_steps)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_prepare(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_prepare(): context is not sharded.");
    const char *ge = getenv("LJMD_SLAB_GRAPH");
    if (max_steps < 2 || max_steps > 32 || (ge && atoi(ge) == 0) || c->graph_slab[max_steps]) return LJMD_OK;
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    NlistWork &w = c->nw;
    const int cur_save = w.cur_host;
    int rc = LJMD_OK;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        c->capturing = true;
        for (int s = 0; s < max_steps && rc == LJMD_OK; ++s) {
            rc = nlist_enqueue_step_kernel(c);
#if !LJMD_NL_PUBLISH_IN_KERNEL
            if (rc == LJMD_OK) {
                const NlArgs a = make_args(c, 0);
                nl_publish_kernel<<<1, 32, 0, c->stream>>>(a);
            }
#endif
        }
        c->capturing = false;
        const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        if (e2 != cudaSuccess) e = e2;
    }
    w.cur_host = cur_save; // capture executes nothing
    if (e == cudaSuccess && rc == LJMD_OK) e = cudaGraphInstantiate(&c->graph_slab[max_steps], g, 0);
    if (g) cudaGraphDestroy(g);
    if (prev != c->device) cudaSetDevice(prev);
    if (rc != LJMD_OK) return rc;
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_prepare(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// Total momentum (sum of v over the atoms the last step / force launch of this context handled).
// Newton's third law makes it an invariant of the dynamics; a sharded run sums the ranks' shares.
extern "C" int ljmd_ctx_nlist_momentum(ljmd_ctx *c, double *p3)
{
    clear_error();
    if (!c || !p3) return set_error(LJMD_ERR_NULL, "ljmd_ctx_nlist_momentum(): NULL argument.");
    if (c->method != LJMD_METHOD_NLIST) return set_error(LJMD_ERR_ARG, "ljmd_ctx_nlist_momentum(): neighbour-list method only.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    cudaError_t e = cudaMemcpyAsync(p3, c->nw.mom, sizeof(double) * 3, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_nlist_momentum(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

extern "C" int ljmd_ctx_slab_publish(ljmd_ctx *c)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_publish(): ctx is NULL.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_publish(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const NlArgs a = make_args(c, 0);
    nl_publish_kernel<<<1, 32, 0, c->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_publish(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

__global__ void nl_poison_kernel(NlArgs a)
{
    if (threadIdx.x != 0) return;
    __threadfence_system();
    for (int q = 0; q < a.world; ++q)
        for (int par = 0; par < 2; ++par) nl_st_release_sys(a.peer_flags[q] + par * LJMD_MAX_RANKS + a.rank, NL_FLAG_POISON);
}

// A rank that cannot go on (an error return from ljmd_ctx_slab_run / _rebuild) tells its peers: its flag reads
// "every step done, skin violated", so the peers' step kernels halt at once instead of waiting for it.
extern "C" int ljmd_ctx_slab_poison(ljmd_ctx *c)
{
    if (!c || !c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_poison(): context is not sharded.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const NlArgs a = make_args(c, 0);
    nl_poison_kernel<<<1, 32, 0, c->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_poison(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// Debug aid (LJMD_SLAB_TIMELINE=1 at context creation): per sharded step since the last call, the device
// globaltimer (ns) of {first CTA start, last CTA past the flag wait, last CTA end, flag published}; re-arms.
extern "C" int ljmd_ctx_debug_slab_timeline(ljmd_ctx *c, uint64_t *out, int32_t max_steps)
{
    if (!c || !out || !c->nw.dbg_time) return 0;
    const int m = max_steps < 1024 ? max_steps : 1024;
    cudaStreamSynchronize(c->stream);
    if (cudaMemcpy(out, c->nw.dbg_time, sizeof(unsigned long long) * 4 * (size_t)m, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    std::vector<unsigned long long> init(4096, 0ull);
    for (int k = 0; k < 1024; ++k) init[4 * k] = ~0ull; // slot 0 is a minimum
    cudaMemcpy(c->nw.dbg_time, init.data(), sizeof(unsigned long long) * 4096, cudaMemcpyHostToDevice);
    return m;
}

// Debug aid: XOR checksums of {next positions, v, a, packed next positions} over ALL slots.
extern "C" int ljmd_ctx_debug_checksum(ljmd_ctx *c, uint64_t *out4)
{
    clear_error();
    if (!c || !out4) return set_error(LJMD_ERR_NULL, "ljmd_ctx_debug_checksum(): NULL argument.");
    if (c->method != LJMD_METHOD_NLIST) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_debug_checksum(): neighbour-list method only.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 32);
    if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, 32, c->stream);
    if (e == cudaSuccess) {
        const NlArgs a = make_args(c, 0);
        nl_checksum_kernel<<<(c->sc.n + 255) / 256, 256, 0, c->stream>>>(a, d);
        e = cudaMemcpyAsync(out4, d, 32, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_debug_checksum(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}

// r(t+dt) of ALL atoms from the complete state (start of a sharded run), verlet.f90:58-63.
extern "C" int ljmd_ctx_slab_drift(ljmd_ctx *c)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_drift(): ctx is NULL.");
    if (c->method != LJMD_METHOD_NLIST) return set_error(LJMD_ERR_UNSUPPORTED, "ljmd_ctx_slab_drift(): needs the neighbour-list method.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    const int rc = nlist_enqueue_drift(c);
    if (prev != c->device) cudaSetDevice(prev);
    return rc;
}

// Up to max_steps sharded velocity-Verlet steps (the step kernels are launched back to back; once
// some rank reports a skin violation the remaining ones are no-ops on every rank).  Returns the
// number of steps actually done and whether the run halted (=> gather + ljmd_ctx_slab_rebuild).
// obs4 (may be NULL) = this rank's share of {epot, ekin, d_epot, dd_epot} of the last step done,
// without tail corrections.
extern "C" int ljmd_ctx_slab_run(ljmd_ctx *c, int32_t max_steps, int32_t *steps_done, int32_t *halted, double *obs4)
{
    clear_error();
    if (!c || !steps_done || !halted) return set_error(LJMD_ERR_NULL, "ljmd_ctx_slab_run(): NULL argument.");
    if (!c->nw.sh_on) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_run(): context is not sharded.");
    if (max_steps < 0) return set_error(LJMD_ERR_ARG, "ljmd_ctx_slab_run(): max_steps must be >= 0.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    NlistWork &w = c->nw;
    int rc = LJMD_OK;
    int *stage = reinterpret_cast<int *>(reinterpret_cast<char *>(c->h_obs) + sizeof(double) * 4 * (size_t)c->R);
    stage[0] = max_steps; stage[1] = 0; // steps_left, steps_done
    cudaError_t e = cudaMemcpyAsync(reinterpret_cast<char *>(w.ctl) + offsetof(NlCtl, steps_left), stage, 2 * sizeof(int),
                                    cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_s0, c->stream);
    // Every step kernel is followed by the 1-thread publish kernel: the kernel boundary guarantees
    // that all halo stores of the step are performed system-wide before its flag becomes visible.
    // (The rare wrong-velocity defect once blamed on adjacent step kernels had another cause -- every
    // thread tested ctl->halted itself -- fixed in the sharded prologue of nl_step_kernel, DESIGN.md 6.
    // LJMD_NL_PUBLISH_IN_KERNEL=1 builds publish from the step kernel's last CTA instead.)
    // The launches of a chunk are replayed as ONE CUDA graph (cached per chunk length): no per-launch host work,
    // shorter gaps between dependent kernels.  LJMD_SLAB_GRAPH=0 issues them one by one.
    const char *ge = getenv("LJMD_SLAB_GRAPH");
    const bool use_graph = max_steps >= 2 && max_steps <= 32 && !(ge && atoi(ge) == 0);
    auto enqueue_chunk = [&]() {
        for (int s = 0; s < max_steps && rc == LJMD_OK; ++s) {
            rc = nlist_enqueue_step_kernel(c);
#if !LJMD_NL_PUBLISH_IN_KERNEL
            if (rc == LJMD_OK) {
                const NlArgs a = make_args(c, 0);
                nl_publish_kernel<<<1, 32, 0, c->stream>>>(a);
                if (!c->capturing) c->launches++;
            }
#endif
        }
    };
    if (use_graph && e == cudaSuccess) {
        if (!c->graph_slab[max_steps]) { // normally captured ahead of time by ljmd_ctx_slab_prepare
            const int cur_save = w.cur_host;
            cudaGraph_t g = nullptr;
            e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                c->capturing = true;
                enqueue_chunk();
                c->capturing = false;
                const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
                if (e2 != cudaSuccess) e = e2;
            }
            w.cur_host = cur_save; // capture executes nothing
            if (e == cudaSuccess && rc == LJMD_OK) e = cudaGraphInstantiate(&c->graph_slab[max_steps], g, 0);
            if (g) cudaGraphDestroy(g);
        }
        if (e == cudaSuccess && rc == LJMD_OK) {
            // the timed region starts after the (one-off) capture
            e = cudaEventRecord(c->ev_s0, c->stream);
            if (e == cudaSuccess) e = cudaGraphLaunch(c->graph_slab[max_steps], c->stream);
            if (max_steps & 1) w.cur_host ^= 1; // what max_steps enqueue calls would have done to the host mirror
#if LJMD_NL_PUBLISH_IN_KERNEL
            c->launches += max_steps;
#else
            c->launches += 2 * (int64_t)max_steps;
#endif
        }
    } else if (e == cudaSuccess) {
        enqueue_chunk();
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_s1, c->stream);
    NlCtl h;
    memset(&h, 0, sizeof(h));
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, w.ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_obs, c->fw.obs, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev_s0, c->ev_s1);
        c->last_step_ms = ms;
    }
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_slab_run(): %s", cudaGetErrorString(e));
    if (rc != LJMD_OK) return rc;
    w.rebuilds_host = h.rebuilds;
    // the host mirror of cur follows the steps that really ran
    if (((max_steps - h.steps_done) & 1) != 0) w.cur_host ^= 1;
    *steps_done = h.steps_done;
    *halted = h.halted;
    if (w.dbg_exec) {
        static int hcount[4096];
        cudaMemcpy(hcount, w.dbg_exec, sizeof(hcount), cudaMemcpyDeviceToHost);
        const int nact_expect = (h.seq > 0) ? hcount[0] : 0; (void)nact_expect;
        for (int q = 0; q < 4096 && q < h.seq + 2; ++q)
            if (hcount[q] != 0 && (hcount[q] < w.own_blocks * 3 / 4 - 2000 || hcount[q] > w.nblk * 4))
                fprintf(stderr, "[ljmd slab debug] rank %d: step seq %d executed by %d CTAs\n", w.rank, q, hcount[q]);
        if (getenv("LJMD_SLAB_DEBUG_DUMP")) {
            fprintf(stderr, "[ljmd slab debug] exec counts up to seq %d:", h.seq);
            for (int q = 0; q <= h.seq && q < 4096; ++q) fprintf(stderr, " %d", hcount[q]);
            fprintf(stderr, "\n");
        }
    }
    if (h.pad_ != 0) fprintf(stderr, "[ljmd slab debug] rank %d: control-block check failed, mask %u (1 = ticket not re-armed, 2 = cur changed under a running CTA, 4 = seq changed)\n", w.rank, h.pad_);
    if (h.overflow > 0) return nlist_check_overflow(c);
    if (obs4) for (int k = 0; k < 4; ++k) obs4[k] = c->h_obs[k];
    return LJMD_OK;
}


extern "C" int ljmd_ctx_nlist_info(ljmd_ctx *c, double *skin, int32_t *cells_per_edge, int32_t *capacity,
                                   int64_t *rebuilds, int64_t *entries)
{
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_nlist_info(): ctx is NULL.");
    if (c->method != LJMD_METHOD_NLIST)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_nlist_info(): context does not use the neighbour list.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    int rc = nlist_check_overflow(c);
    long long total = 0;
    if (rc == LJMD_OK && entries) {
        std::vector<int> cnt(c->sc.n);
        cudaError_t e = cudaMemcpy(cnt.data(), c->nw.nb_cnt, sizeof(int) * (size_t)c->sc.n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(LJMD_ERR_CUDA, "ljmd_ctx_nlist_info(): %s", cudaGetErrorString(e));
        int kb = 0, ke = c->sc.n;
        if (rc == LJMD_OK && c->nw.sh_on) { // sharded: only the owned slot range has current lists
            int32_t bounds[LJMD_MAX_RANKS + 1];
            rc = nlist_slab_bounds_host(c, bounds);
            if (rc == LJMD_OK) { kb = bounds[c->nw.rank]; ke = bounds[c->nw.rank + 1]; }
        }
        for (int k = kb; k < ke; ++k) total += cnt[k];
    }
    if (prev != c->device) cudaSetDevice(prev);
    if (rc != LJMD_OK) return rc;
    if (skin) *skin = c->nw.skin;
    if (cells_per_edge) *cells_per_edge = c->nw.nc;
    if (capacity) *capacity = c->nw.maxnb;
    if (rebuilds) *rebuilds = c->nw.rebuilds_host;
    if (entries) *entries = total;
    return LJMD_OK;
}

// Neighbour list of the last build in ORIGINAL atom numbering: counts[n]; the neighbours of atom i,
// ascending, at nbr[i*cap .. i*cap + counts[i]).
extern "C" int ljmd_ctx_download_nlist(ljmd_ctx *c, int32_t *counts, int32_t *nbr, int32_t cap)
{
    clear_error();
    if (!c || !counts || !nbr) return set_error(LJMD_ERR_NULL, "ljmd_ctx_download_nlist(): NULL argument.");
    if (c->method != LJMD_METHOD_NLIST)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_download_nlist(): context does not use the neighbour list.");
    const NlistWork &w = c->nw;
    const int n = c->sc.n;
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    std::vector<int> h_orig(n), h_cnt(n), h_nbr(w.stride * (size_t)w.maxnb);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(h_orig.data(), w.orig, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_cnt.data(), w.nb_cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_nbr.data(), w.nbr, sizeof(int) * h_nbr.size(), cudaMemcpyDeviceToHost);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_download_nlist(): %s", cudaGetErrorString(e));
    for (int k = 0; k < n; ++k) {
        const int o = h_orig[k], m = h_cnt[k];
        if (m > cap) return set_error(LJMD_ERR_ARG, "ljmd_ctx_download_nlist(): cap %d < %d neighbours.", cap, m);
        counts[o] = m;
        int32_t *dst = nbr + (size_t)o * cap;
        for (int q = 0; q < m; ++q) dst[q] = h_orig[h_nbr[nl_entry(w.stride, k, q)]];
        std::sort(dst, dst + m);
    }
    return LJMD_OK;
}
