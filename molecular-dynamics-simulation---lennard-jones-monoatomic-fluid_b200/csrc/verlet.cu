// verlet.cu -- O(N) kernels of the velocity-Verlet step and the deterministic scalar
// reductions.  Compiled with -fmad=false: every operation below is rounded exactly as the
// reference's gfortran -O2 build rounds it (no contraction), in the reference's order.
//
// Reference (file:line relative to /root/reference/scripts/):
//   physics/verlet.f90:58-60   drift        r = (r + v*dt) + a*dt_square_half
//   physics/geometry_pbc.f90:53-57 wrap     x = x - L*floor(x*invL)
//   physics/verlet.f90:72-74   half kick    v = v + a*dt_half
//   physics/verlet.f90:86-88   half kick    v = v + a_new*dt_half
//   physics/verlet.f90:93-95   E_kin        0.5*(sum(vx^2)+sum(vy^2)+sum(vz^2))
//   physics/lj_potential_energy.f90:188-223 prefactors 4/24 and tail corrections
//   md_simulation_program.f90:339-353       unwrapped update ru += MIC(r_new - r_old)
#include "ljmd_internal.h"
#include "reduce.cuh"

namespace ljmd {

// ------------------------------------------------------------------------------------------
// drift + wrap + first half kick (+ unwrapped-coordinate update).  One thread per atom.
// HBM traffic per atom: read r,v,a (72 B) + write r,v (48 B) [+ ru read/write 48 B].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) verlet_pre_kernel(SimConst sc, StatePtrs st, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double L = sc.L;
    // geometry_pbc.f90:51 recomputes invL = 1/L locally; same value as params%inv_box_length.
    const double invL = sc.invL;
    const double x0 = st.rx[i], y0 = st.ry[i], z0 = st.rz[i];
    const double vx = st.vx[i], vy = st.vy[i], vz = st.vz[i];
    const double ax = st.ax[i], ay = st.ay[i], az = st.az[i];
    double x = (x0 + vx * sc.dt) + ax * sc.dt_square_half; // verlet.f90:58
    double y = (y0 + vy * sc.dt) + ay * sc.dt_square_half;
    double z = (z0 + vz * sc.dt) + az * sc.dt_square_half;
    x = x - L * floor(x * invL); // geometry_pbc.f90:54-56
    y = y - L * floor(y * invL);
    z = z - L * floor(z * invL);
    st.rx[i] = x; st.ry[i] = y; st.rz[i] = z;
    st.vx[i] = vx + ax * sc.dt_half; // verlet.f90:72-74
    st.vy[i] = vy + ay * sc.dt_half;
    st.vz[i] = vz + az * sc.dt_half;
    if (st.rux) { // md_simulation_program.f90:341-351 (dnint = round half away from zero)
        double dx = x - x0, dy = y - y0, dz = z - z0;
        dx = dx - L * round(dx * invL);
        dy = dy - L * round(dy * invL);
        dz = dz - L * round(dz * invL);
        st.rux[i] = st.rux[i] + dx;
        st.ruy[i] = st.ruy[i] + dy;
        st.ruz[i] = st.ruz[i] + dz;
    }
}

int launch_verlet_pre(ljmd_ctx *c)
{
    const size_t total = c->nR;
    const int block = 256;
    const unsigned grid = (unsigned)((total + block - 1) / block);
    verlet_pre_kernel<<<grid, block, 0, c->stream>>>(c->sc, c->st, total);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// ------------------------------------------------------------------------------------------
// Finalize: a_i = 24 * (sum of the force partials of atom i, fixed order), optional second
// half kick and E_kin partials, then the LAST CTA of each replica reduces every scalar partial
// in a fixed order and writes {epot, ekin, d_epot, dd_epot}.
// Which CTA is last varies run to run, but what it computes does not.
// ------------------------------------------------------------------------------------------
struct FinArgs {
    SimConst sc;
    StatePtrs st;
    const double *fpart; // [R][n_splits][3][n]
    int n_splits;
    const double *spart; // [R][n_spart][3]
    int n_spart;
    double pair_weight;  // 0.5 when spart sums ORDERED pairs, 1.0 for unordered
    double *kpart;       // [R][gridDim.x][3]
    unsigned *counter;   // [R]
    double *obs;         // [R][4]
    StepCtl *ctl;
    int kick;
};

__global__ void __launch_bounds__(256) finalize_kernel(FinArgs a)
{
    __shared__ double scratch[3 * 8];
    __shared__ bool is_last;
    const int n = a.sc.n;
    const int rep = blockIdx.y;
    const size_t off = (size_t)rep * n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;

    double k[3] = {0.0, 0.0, 0.0};
    if (i < n) {
        const double *fp = a.fpart + (size_t)rep * a.n_splits * 3 * n;
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for (int s = 0; s < a.n_splits; ++s) { // fixed order
            fx += fp[((size_t)s * 3 + 0) * n + i];
            fy += fp[((size_t)s * 3 + 1) * n + i];
            fz += fp[((size_t)s * 3 + 2) * n + i];
        }
        const double ax = 24.0 * fx, ay = 24.0 * fy, az = 24.0 * fz; // lj_potential_energy.f90:189-191
        a.st.ax[off + i] = ax; a.st.ay[off + i] = ay; a.st.az[off + i] = az;
        double vx = a.st.vx[off + i], vy = a.st.vy[off + i], vz = a.st.vz[off + i];
        if (a.kick) { // verlet.f90:86-88
            vx = vx + ax * a.sc.dt_half;
            vy = vy + ay * a.sc.dt_half;
            vz = vz + az * a.sc.dt_half;
            a.st.vx[off + i] = vx; a.st.vy[off + i] = vy; a.st.vz[off + i] = vz;
        }
        k[0] = vx * vx; k[1] = vy * vy; k[2] = vz * vz; // verlet.f90:93-95
    }
    block_sum<3>(k, scratch);
    if (threadIdx.x == 0) {
        double *kp = a.kpart + ((size_t)rep * gridDim.x + blockIdx.x) * 3;
        kp[0] = k[0]; kp[1] = k[1]; kp[2] = k[2];
        __threadfence();
        const unsigned ticket = atomicAdd(&a.counter[rep], 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // Fixed-order reduction of all partials of this replica.
    double s[3] = {0.0, 0.0, 0.0};
    const double *sp = a.spart + (size_t)rep * a.n_spart * 3;
    for (int b = threadIdx.x; b < a.n_spart; b += blockDim.x) {
        s[0] += __ldcg(sp + 3 * b + 0);
        s[1] += __ldcg(sp + 3 * b + 1);
        s[2] += __ldcg(sp + 3 * b + 2);
    }
    block_sum<3>(s, scratch);
    double kk[3] = {0.0, 0.0, 0.0};
    const double *kp = a.kpart + (size_t)rep * gridDim.x * 3;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        kk[0] += __ldcg(kp + 3 * b + 0);
        kk[1] += __ldcg(kp + 3 * b + 1);
        kk[2] += __ldcg(kp + 3 * b + 2);
    }
    block_sum<3>(kk, scratch);
    if (threadIdx.x == 0) {
        const double w = a.pair_weight;
        double o[4];
        o[0] = 4.0 * (w * s[0]) + a.sc.tail_u;    // lj_potential_energy.f90:188,221
        o[1] = 0.5 * (kk[0] + kk[1] + kk[2]);     // verlet.f90:93-95
        o[2] = 24.0 * (w * s[1]) + a.sc.tail_du;  // :192,222
        o[3] = 24.0 * (w * s[2]) + a.sc.tail_ddu; // :193,223
        double *ob = a.obs + 4 * (size_t)rep;
        ob[0] = o[0]; ob[1] = o[1]; ob[2] = o[2]; ob[3] = o[3];
        StepCtl *ctl = a.ctl;
        if (a.kick) {
            const long long step = ctl->step;
            if (ctl->series && step < ctl->series_cap) {
                double *sr = ctl->series + ((size_t)step * gridDim.y + rep) * 4;
                sr[0] = o[0]; sr[1] = o[1]; sr[2] = o[2]; sr[3] = o[3];
            }
            // the replica with the highest index that finishes bumps the step: use a ticket
            // so that exactly one CTA per step does it after all replicas wrote their row.
            __threadfence();
            const unsigned t2 = atomicAdd(&a.counter[gridDim.y], 1u);
            if (t2 == gridDim.y - 1) {
                a.counter[gridDim.y] = 0;
                ctl->step = step + 1;
            }
        }
        a.counter[rep] = 0; // re-arm for the next launch
    }
}

int launch_finalize(ljmd_ctx *c, bool kick)
{
    FinArgs a;
    a.sc = c->sc;
    a.st = c->st;
    a.fpart = c->pw.fpart;
    a.n_splits = c->pw.n_splits;
    a.spart = c->pw.spart;
    a.n_spart = c->pw.n_itiles * c->pw.n_splits;
    a.pair_weight = (c->pw.variant == 0) ? 0.5 : 1.0;
    a.kpart = c->fw.kpart;
    a.counter = c->fw.counter;
    a.obs = c->fw.obs;
    a.ctl = c->fw.ctl;
    a.kick = kick ? 1 : 0;
    dim3 grid(c->fw.nblk, c->R);
    finalize_kernel<<<grid, c->fw.block, 0, c->stream>>>(a);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// ------------------------------------------------------------------------------------------
// Driver-side kinetic energy 0.5*sum_i (vx^2+vy^2+vz^2), md_simulation_program.f90:238-240.
// One CTA per replica, fixed-order strided accumulation + tree.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ekin_driver_kernel(int n, const double *vx,
                                                          const double *vy, const double *vz,
                                                          double *out)
{
    __shared__ double scratch[8];
    const size_t off = (size_t)blockIdx.x * n;
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = vx[off + i], y = vy[off + i], z = vz[off + i];
        s[0] += (x * x + y * y) + z * z;
    }
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = 0.5 * s[0];
}

int launch_ekin_driver(ljmd_ctx *c, double *d_out)
{
    ekin_driver_kernel<<<c->R, 256, 0, c->stream>>>(c->sc.n, c->st.vx, c->st.vy, c->st.vz, d_out);
    c->launches++;
    LJMD_CUDA_CHECK(cudaGetLastError());
    return LJMD_OK;
}

// ------------------------------------------------------------------------------------------
// FP64 peak: 8 independent register-resident DFMA chains per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-9, x2 = x0 + 2e-9, x3 = x0 + 3e-9;
    double x4 = x0 + 4e-9, x5 = x0 + 5e-9, x6 = x0 + 6e-9, x7 = x0 + 7e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s; // never true; keeps the chains alive
}

int measure_fp64_peak(int device, double *tflops)
{
    int prev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&prev));
    if (device >= 0) LJMD_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    LJMD_CUDA_CHECK(cudaGetDevice(&dev));
    LJMD_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    double *d = nullptr;
    LJMD_CUDA_CHECK(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    LJMD_CUDA_CHECK(cudaEventCreate(&e0));
    LJMD_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 2048, block = 256, grid = prop.multiProcessorCount * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        LJMD_CUDA_CHECK(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        LJMD_CUDA_CHECK(cudaEventRecord(e1));
        LJMD_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        LJMD_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)block * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (device >= 0) cudaSetDevice(prev);
    *tflops = best;
    return LJMD_OK;
}

} // namespace ljmd
