// lj_cells.cu -- cell-list force path for large N (no reference counterpart: the reference is
// all-pairs only, /root/reference/scripts/physics/lj_potential_energy.f90:109-183; the pair
// arithmetic and the minimum image are the same as there).
//
// Per force evaluation:
//   1. cell_count   : cell id of every atom (c = clamp(floor(x * nc/L)), id = (cx*nc+cy)*nc+cz)
//                     and a histogram of the cells (integer atomics: order-independent result).
//   2. cell_scan    : exclusive prefix sum of the histogram -> cell_start.
//   3. cell_fill    : scatter atom indices into their cell's segment (arbitrary order) ...
//   4. cell_sort    : ... then every segment is sorted by atom index, so that the result is the
//                     STABLE counting sort (ascending original index inside a cell), identical
//                     run to run and bit-exact against the CPU definition; positions are gathered
//                     into cell-sorted SoA arrays.
//   5. cells_force  : one warp per (column (cx,cy), group of G cells along z).  Cells of one
//                     column are contiguous in the sorted order, so the 27-cell stencil of the
//                     group is 9 contiguous strips (two pieces where z wraps).  Each lane owns one
//                     i-atom and walks the strips, staged 32 atoms at a time in shared memory.
//                     Ordered pairs (every thread accumulates only its own atom => no atomics,
//                     fixed summation order), energy-like sums carry weight 1/2.
//   The fused finalize kernel (verlet.cu) gathers the sorted forces back through slot_of_atom.
#include <cmath>

#include "ljmd_internal.h"
#include "lj_pair.cuh"
#include "reduce.cuh"

namespace ljmd {

int cells_per_edge_host(const ljmd_params &p)
{
    // Largest nc with cell edge L/nc >= rc, backed off by one when the edge is within 1e-12
    // relative of rc so that a rounding in floor(x*scale) can never separate an in-cutoff pair by
    // two cells; at most 128 cells per edge.
    volatile double ratio = p.box_length / p.rc;
    int nc = (int)floor(ratio);
    if (nc > 0 && (double)nc * p.rc * (1.0 + 1e-12) > p.box_length) nc -= 1;
    if (nc > 128) nc = 128;
    return nc;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double x, double scale, int nc, double L, double invL)
{
    // out-of-box coordinates (unwrapped input, geometry_pbc.f90:18-19) are wrapped for binning
    if (x < 0.0 || x > L) x = x - L * floor(x * invL);
    int c = (int)floor(x * scale); // one rounding, as on the host
    c = max(c, 0);
    return min(c, nc - 1);         // wrapped x may equal L exactly (geometry_pbc.f90:53-57)
}

__global__ void __launch_bounds__(256) cell_count_kernel(int n, int nc, double scale, double L, double invL,
                                                         const double *rx, const double *ry, const double *rz,
                                                         int *cell_of_atom, int *count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coord(rx[i], scale, nc, L, invL), cy = cell_coord(ry[i], scale, nc, L, invL),
              cz = cell_coord(rz[i], scale, nc, L, invL);
    const int c = (cx * nc + cy) * nc + cz;
    cell_of_atom[i] = c;
    atomicAdd(&count[c], 1);
}

// Exclusive scan of count[0..ncells) into start[0..ncells]; one CTA, running carry.
__global__ void __launch_bounds__(1024) cell_scan_kernel(int ncells, const int *count, int *start, int *fill)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < ncells; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < ncells ? count[idx] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
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
        const int prefix = carry + (warp > 0 ? wsum[warp - 1] : 0) + x - v; // exclusive
        if (idx < ncells) { start[idx] = prefix; fill[idx] = 0; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[ncells] = carry;
}

__global__ void __launch_bounds__(256) cell_fill_kernel(int n, const int *cell_of_atom, const int *start, int *fill,
                                                        int *order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of_atom[i];
    order[start[c] + atomicAdd(&fill[c], 1)] = i;
}

// One thread per cell: insertion sort of the (short) segment by atom index, then gather.
__global__ void __launch_bounds__(128) cell_sort_kernel(int ncells, int n, const int *start, int *order,
                                                        int *slot_of_atom, const double *rx, const double *ry,
                                                        const double *rz, double *sx, double *sy, double *sz)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int b = start[c], e = start[c + 1];
    for (int k = b + 1; k < e; ++k) {
        const int v = order[k];
        int m = k - 1;
        while (m >= b && order[m] > v) { order[m + 1] = order[m]; --m; }
        order[m + 1] = v;
    }
    for (int k = b; k < e; ++k) {
        const int i = order[k];
        slot_of_atom[i] = k;
        sx[k] = rx[i]; sy[k] = ry[i]; sz[k] = rz[i];
    }
}

// ------------------------------------------------------------------------------------------
struct CellForceArgs {
    SimConst sc;
    int nc, G, nzg, nunits;
    const int *start;
    const int *order;   // sorted slot -> original atom index
    int own_begin, own_end; // only atoms with original index in [own_begin, own_end) get forces
    const double *sx, *sy, *sz;
    double *fsorted; // [3][n]
    double *spart;   // [gridDim.x][3]
    unsigned long long *pair_counts;
};

constexpr int CF_WARPS = 4;

template <bool COUNT>
__global__ void __launch_bounds__(CF_WARPS * 32) cells_force_kernel(CellForceArgs a)
{
    __shared__ double sj[CF_WARPS][3][32];
    __shared__ double scratch[3 * CF_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.sc.n, nc = a.nc;
    const double L = a.sc.L, invL = a.sc.invL, rc2 = a.sc.rc2;
    double u = 0.0, du = 0.0, ddu = 0.0;
    unsigned long long ncand = 0, ncut = 0;

    const int unit = blockIdx.x * CF_WARPS + warp;
    if (unit < a.nunits) {
        const int zg = unit % a.nzg, col = unit / a.nzg;
        const int cy = col % nc, cx = col / nc;
        const int cz0 = zg * a.G, cz1 = min(nc, cz0 + a.G) - 1;
        const int colbase = (cx * nc + cy) * nc;
        const int ib = a.start[colbase + cz0], ie = a.start[colbase + cz1 + 1];
        // z window of the neighbours: cells cz0-1 .. cz1+1 with periodic wrap (whole column when
        // the window covers it)
        int zlo = cz0 - 1, zhi = cz1 + 1;
        const bool whole = (zhi - zlo + 1 >= nc);
        for (int i0 = ib; i0 < ie; i0 += 32) {
            const int k = i0 + lane;
            bool vi = k < ie;
            if (vi) { const int io = a.order[k]; vi = (io >= a.own_begin) && (io < a.own_end); }
            if (!__any_sync(0xffffffffu, vi)) continue; // nothing owned in this batch (sharded runs)
            const double xi = vi ? a.sx[k] : 0.0, yi = vi ? a.sy[k] : 0.0, zi = vi ? a.sz[k] : 0.0;
            double fx = 0.0, fy = 0.0, fz = 0.0;
            for (int ox = -1; ox <= 1; ++ox) {
                const int nx = (cx + ox + nc) % nc;
                for (int oy = -1; oy <= 1; ++oy) {
                    const int ny = (cy + oy + nc) % nc;
                    const int nb = (nx * nc + ny) * nc;
                    // up to two contiguous pieces of the neighbour column
                    int sb[2], se[2], npieces;
                    if (whole) {
                        sb[0] = a.start[nb]; se[0] = a.start[nb + nc]; npieces = 1;
                    } else if (zlo < 0) {
                        sb[0] = a.start[nb + zlo + nc]; se[0] = a.start[nb + nc];
                        sb[1] = a.start[nb]; se[1] = a.start[nb + zhi + 1]; npieces = 2;
                    } else if (zhi >= nc) {
                        sb[0] = a.start[nb + zlo]; se[0] = a.start[nb + nc];
                        sb[1] = a.start[nb]; se[1] = a.start[nb + zhi - nc + 1]; npieces = 2;
                    } else {
                        sb[0] = a.start[nb + zlo]; se[0] = a.start[nb + zhi + 1]; npieces = 1;
                    }
                    for (int pc = 0; pc < npieces; ++pc) {
                        for (int j0 = sb[pc]; j0 < se[pc]; j0 += 32) {
                            const int cnt = min(32, se[pc] - j0);
                            __syncwarp();
                            if (lane < cnt) {
                                sj[warp][0][lane] = a.sx[j0 + lane];
                                sj[warp][1][lane] = a.sy[j0 + lane];
                                sj[warp][2][lane] = a.sz[j0 + lane];
                            }
                            __syncwarp();
                            for (int jj = 0; jj < cnt; ++jj) {
                                const double dx = mic(xi - sj[warp][0][jj], L, invL);
                                const double dy = mic(yi - sj[warp][1][jj], L, invL);
                                const double dz = mic(zi - sj[warp][2][jj], L, invL);
                                const double r2 = dist2(dx, dy, dz);
                                const bool other = vi && (j0 + jj != k);
                                if (COUNT && other) ++ncand;
                                if (r2 < rc2 && other) {
                                    const PairTerms t = pair_terms(dx, dy, dz, r2);
                                    fx += t.fx; fy += t.fy; fz += t.fz;
                                    u += t.u; du += t.du; ddu += t.ddu;
                                    if (COUNT) ++ncut;
                                }
                            }
                        }
                    }
                }
            }
            if (vi) {
                a.fsorted[k] = fx; a.fsorted[(size_t)n + k] = fy; a.fsorted[2 * (size_t)n + k] = fz;
            }
        }
    }
    double s[3] = {u, du, ddu};
    block_sum<3>(s, scratch);
    if (threadIdx.x == 0) {
        double *sp = a.spart + (size_t)blockIdx.x * 3;
        sp[0] = s[0]; sp[1] = s[1]; sp[2] = s[2];
    }
    if (COUNT) {
        // integer counters: order-independent
        for (int o = 16; o > 0; o >>= 1) {
            ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
            ncut += __shfl_xor_sync(0xffffffffu, ncut, o);
        }
        if (lane == 0 && a.pair_counts) { atomicAdd(&a.pair_counts[0], ncand); atomicAdd(&a.pair_counts[1], ncut); }
    }
}

// ------------------------------------------------------------------------------------------
int cellwork_plan(ljmd_ctx *c)
{
    CellWork &w = c->cw;
    if (c->R != 1) return set_error(LJMD_ERR_UNSUPPORTED, "cell list supports nreplicas == 1 only");
    const int n = c->sc.n;
    w.nc = cells_per_edge_host(c->p);
    if (w.nc < 3) return set_error(LJMD_ERR_UNSUPPORTED, "cell list needs floor(L/rc) >= 3 (got %d)", w.nc);
    w.ncells = w.nc * w.nc * w.nc;
    w.scale = (double)w.nc * c->p.inv_box_length;
    w.G = 2;
    w.nzg = (w.nc + w.G - 1) / w.G;
    w.nunits = w.nc * w.nc * w.nzg;
    w.nblk_force = (w.nunits + CF_WARPS - 1) / CF_WARPS;
    w.kernels_per_force = 5;
    w.pair_weight = 0.5;
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_of_atom, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_start, sizeof(int) * ((size_t)w.ncells + 1)));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_fill, sizeof(int) * (size_t)w.ncells));
    LJMD_CUDA_CHECK(cudaMalloc(&w.cell_count, sizeof(int) * (size_t)w.ncells));
    LJMD_CUDA_CHECK(cudaMalloc(&w.order, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.slot_of_atom, sizeof(int) * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.sx, sizeof(double) * 3 * (size_t)n));
    w.sy = w.sx + n;
    w.sz = w.sy + n;
    LJMD_CUDA_CHECK(cudaMalloc(&w.fsorted, sizeof(double) * 3 * (size_t)n));
    LJMD_CUDA_CHECK(cudaMalloc(&w.spart, sizeof(double) * 3 * (size_t)w.nblk_force));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.fsorted, 0, sizeof(double) * 3 * (size_t)n, c->stream));
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.spart, 0, sizeof(double) * 3 * (size_t)w.nblk_force, c->stream));
    return LJMD_OK;
}

void cellwork_free(ljmd_ctx *c)
{
    CellWork &w = c->cw;
    cudaFree(w.cell_of_atom); cudaFree(w.cell_start); cudaFree(w.cell_fill); cudaFree(w.cell_count);
    cudaFree(w.order); cudaFree(w.slot_of_atom); cudaFree(w.sx); cudaFree(w.fsorted); cudaFree(w.spart);
}

int launch_cells_force(ljmd_ctx *c, int buf)
{
    CellWork &w = c->cw;
    const int n = c->sc.n;
    const double *px = c->rbuf[buf], *py = px + c->nR, *pz = py + c->nR;
    LJMD_CUDA_CHECK(cudaMemsetAsync(w.cell_count, 0, sizeof(int) * (size_t)w.ncells, c->stream));
    cell_count_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, w.nc, w.scale, c->sc.L, c->sc.invL, px, py, pz, w.cell_of_atom,
                                                               w.cell_count);
    cell_scan_kernel<<<1, 1024, 0, c->stream>>>(w.ncells, w.cell_count, w.cell_start, w.cell_fill);
    cell_fill_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, w.cell_of_atom, w.cell_start, w.cell_fill, w.order);
    cell_sort_kernel<<<(w.ncells + 127) / 128, 128, 0, c->stream>>>(w.ncells, n, w.cell_start, w.order, w.slot_of_atom,
                                                                   px, py, pz, w.sx, w.sy, w.sz);
    CellForceArgs a;
    a.sc = c->sc;
    a.nc = w.nc; a.G = w.G; a.nzg = w.nzg; a.nunits = w.nunits;
    a.start = w.cell_start;
    a.order = w.order;
    a.own_begin = c->sharded ? c->shard_begin : 0;
    a.own_end = c->sharded ? c->shard_end : n;
    a.sx = w.sx; a.sy = w.sy; a.sz = w.sz;
    a.fsorted = w.fsorted;
    a.spart = w.spart;
    a.pair_counts = c->fw.pair_counts;
    if (a.pair_counts) cells_force_kernel<true><<<w.nblk_force, CF_WARPS * 32, 0, c->stream>>>(a);
    else cells_force_kernel<false><<<w.nblk_force, CF_WARPS * 32, 0, c->stream>>>(a);
    if (!c->capturing) c->launches += w.kernels_per_force;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

} // namespace ljmd

extern "C" int ljmd_ctx_download_cells(ljmd_ctx *c, int32_t *cell_of_atom, int32_t *cell_start, int32_t *order)
{
    using namespace ljmd;
    clear_error();
    if (!c) return set_error(LJMD_ERR_NULL, "ljmd_ctx_download_cells(): ctx is NULL.");
    if (c->method != LJMD_METHOD_CELLS && c->method != LJMD_METHOD_NLIST)
        return set_error(LJMD_ERR_ARG, "ljmd_ctx_download_cells(): context does not use the cell list.");
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (prev != c->device) LJMD_CUDA_CHECK(cudaSetDevice(c->device));
    if (c->method == LJMD_METHOD_NLIST) {
        const int rc = nlist_download_cells(c, cell_of_atom, cell_start, order);
        if (prev != c->device) cudaSetDevice(prev);
        return rc;
    }
    const CellWork &w = c->cw;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && cell_of_atom)
        e = cudaMemcpy(cell_of_atom, w.cell_of_atom, sizeof(int) * (size_t)c->sc.n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cell_start)
        e = cudaMemcpy(cell_start, w.cell_start, sizeof(int) * ((size_t)w.ncells + 1), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && order)
        e = cudaMemcpy(order, w.order, sizeof(int) * (size_t)c->sc.n, cudaMemcpyDeviceToHost);
    if (prev != c->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(LJMD_ERR_CUDA, "ljmd_ctx_download_cells(): %s", cudaGetErrorString(e));
    return LJMD_OK;
}
