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
