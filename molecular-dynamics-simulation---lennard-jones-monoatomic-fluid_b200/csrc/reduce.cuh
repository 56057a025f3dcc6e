// reduce.cuh -- deterministic (fixed-shape) warp/block reductions in FP64.
// No atomics on floating-point data anywhere: results are bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>

namespace ljmd {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sum of `v` over the CTA; every thread of warp 0 returns the total.  `scratch` needs
// (blockDim.x/32) doubles per quantity.  Warp partials are added in warp order.
template <int NQ>
__device__ __forceinline__ void block_sum(double (&v)[NQ], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = warp_sum(v[q]);
    __syncthreads(); // scratch may be reused by the caller
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) scratch[q * nwarp + warp] = v[q];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double s = 0.0;
            for (int w = 0; w < nwarp; ++w) s += scratch[q * nwarp + w];
            v[q] = s;
        }
    }
}

} // namespace ljmd
