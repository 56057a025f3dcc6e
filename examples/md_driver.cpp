// md_driver.cpp -- a C++ host that drives the GPU hot path through the C ABI only (include/ljmd.h),
// following the call sequence of the reference's production program
// (/root/reference/scripts/md_simulation_program.f90:208-391): state in, t=0 forces (:236), then
// total_steps velocity-Verlet steps sampled every output_interval steps after warmup_steps (:361),
// with the unwrapped coordinates tracked on the device (:339-353).  Host arrays are touched only at
// the sampling points; between them the state stays in HBM (ljmd_ctx_step).
//
//   g++ -O2 -Iinclude examples/md_driver.cpp -L<pkg>/lib -lljmd -Wl,-rpath,<pkg>/lib -o md_driver
//   ./md_driver [k=5] [L=10] [rc=4.9] [dt=1e-4] [total_steps=2000] [output_interval=100] [warmup=500]
//
// The initial state here is the FCC lattice of md_initial_config_program.f90:132-187 with zero
// velocities (a host that needs the reference's random velocities uses initial_config.py or its own
// port of base/random_numbers.f90).  Exit status: 0, or the LJMD_ERR_* code with the message the
// reference would have `stop`ped with.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ljmd.h"

static int fail(int code)
{
    std::fprintf(stderr, "md_driver: %s (status %d)\n", ljmd_last_error(), code);
    return code;
}

int main(int argc, char **argv)
{
    const int k = argc > 1 ? std::atoi(argv[1]) : 5;
    const double L = argc > 2 ? std::atof(argv[2]) : 10.0;
    const double rc = argc > 3 ? std::atof(argv[3]) : 4.9;
    const double dt = argc > 4 ? std::atof(argv[4]) : 1e-4;
    const int total_steps = argc > 5 ? std::atoi(argv[5]) : 2000;
    const int output_interval = argc > 6 ? std::atoi(argv[6]) : 100;
    const int warmup_steps = argc > 7 ? std::atoi(argv[7]) : 500;
    const int n = 4 * k * k * k; // base/read_input_files.f90:168

    ljmd_params p;
    int rc_ = ljmd_params_init(&p, n, L, dt, rc, k);
    if (rc_ != LJMD_OK) return fail(rc_);

    // FCC lattice, md_initial_config_program.f90:132-187
    std::vector<double> r[3], v[3], a[3], ru[3];
    for (int c = 0; c < 3; ++c) { r[c].resize(n); v[c].assign(n, 0.0); a[c].assign(n, 0.0); ru[c].resize(n); }
    const double cell = L / k;
    int idx = 0;
    for (int ix = 0; ix < k; ++ix)
        for (int iy = 0; iy < k; ++iy)
            for (int iz = 0; iz < k; ++iz) {
                const double x0 = ix * cell, y0 = iy * cell, z0 = iz * cell, h = 0.5 * cell;
                const double bx[4] = {x0, x0, x0 + h, x0 + h}, by[4] = {y0, y0 + h, y0, y0 + h},
                             bz[4] = {z0, z0 + h, z0 + h, z0};
                for (int b = 0; b < 4; ++b, ++idx) { r[0][idx] = bx[b]; r[1][idx] = by[b]; r[2][idx] = bz[b]; }
            }

    ljmd_options o = {LJMD_METHOD_AUTO, -1, 1, LJMD_FLAG_TRACK_UNWRAPPED};
    ljmd_ctx *ctx = nullptr;
    if ((rc_ = ljmd_ctx_create(&p, &o, &ctx)) != LJMD_OK) return fail(rc_);
    if ((rc_ = ljmd_ctx_upload(ctx, r[0].data(), r[1].data(), r[2].data(), v[0].data(), v[1].data(), v[2].data(),
                               nullptr, nullptr, nullptr)) != LJMD_OK) return fail(rc_);
    double obs3[3], obs4[4];
    if ((rc_ = ljmd_ctx_compute_forces(ctx, obs3)) != LJMD_OK) return fail(rc_); // md_simulation_program.f90:236
    std::printf("# t=0  epot %.12e  d_epot %.12e  dd_epot %.12e\n# time epot ekin etot T P\n", obs3[0], obs3[1], obs3[2]);

    for (int step = output_interval; step <= total_steps; step += output_interval) {
        if ((rc_ = ljmd_ctx_step(ctx, output_interval, obs4)) != LJMD_OK) return fail(rc_);
        if (step > warmup_steps) { // :361
            const double T = 2.0 * obs4[1] / (3.0 * n);                              // stats/md_means.f90:221
            const double P = p.density * T + (-obs4[2]) / (3.0 * p.volume);          // :227, virial = -d_epot
            std::printf("%13.6e %13.6e %13.6e %13.6e %13.6e %13.6e\n", step * dt, obs4[0], obs4[1], obs4[0] + obs4[1], T, P);
        }
    }
    // the snapshot the reference writes into rva.dat (:384-387): r, ru, v, a
    if ((rc_ = ljmd_ctx_download(ctx, r[0].data(), r[1].data(), r[2].data(), v[0].data(), v[1].data(), v[2].data(),
                                 a[0].data(), a[1].data(), a[2].data())) != LJMD_OK) return fail(rc_);
    if ((rc_ = ljmd_ctx_download_unwrapped(ctx, ru[0].data(), ru[1].data(), ru[2].data())) != LJMD_OK) return fail(rc_);
    std::printf("# final r[0] = (%.6f, %.6f, %.6f), launches %lld\n", r[0][0], r[1][0], r[2][0],
                (long long)ljmd_ctx_launch_count(ctx));
    ljmd_ctx_destroy(ctx);
    return 0;
}
