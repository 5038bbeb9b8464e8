// Shared helpers for the draupnir_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define DRP_API extern "C" __attribute__((visibility("default")))

// Return convention of every entry point (include/draupnir_b200.h):
//   0 ok, -k = argument k is invalid, 1000+e = CUDA runtime error e at launch.
static inline int drp_launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < nw ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// fp64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  Lane layout: gid = lane>>2, tig = lane&3;
// a = A[gid][tig], b = B[tig][gid], c0/c1 = C[gid][2*tig], C[gid][2*tig+1].
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// exp for finite arguments without the library routine's special-case handling (clamped to +-700): n = rint(x log2 e),
// two-term Cody-Waite reduction, degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18, observed <= 1 ulp),
// scaling by 2^n through the exponent field.  The fp64 pipe is what bounds the covariance assembly (30 exponentials per
// distance), the categorical kernel (21 per site) and the GRU gates, so the ~30 % fewer instructions show up directly.
__device__ __forceinline__ double drp_exp(double x)
{
    x = fmin(fmax(x, -700.0), 700.0);
    const double t = rint(x * 1.4426950408889634074);
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.0876756987868100e-09);        // 1/12!
    p = fma(p, r, 2.5052108385441720e-08);        // 1/11!
    p = fma(p, r, 2.7557319223985893e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985888e-06);        // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);        // 1/8!
    p = fma(p, r, 1.9841269841269841e-04);        // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);        // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + ((int)t << 20), __double2loint(p));
}
