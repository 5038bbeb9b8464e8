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

// exp for finite arguments without the library routine's special-case handling (clamped to +-700):
// x = (32 m + j) ln2/32 + r with |r| <= ln2/64, exp(x) = 2^m * 2^(j/32) * exp(r); 2^(j/32) from a 32-entry table (L1-resident),
// exp(r) - 1 by a degree-6 polynomial (truncation 1e-18), two-term Cody-Waite reduction; observed error <= 1 ulp.  About
// half the fp64 instructions of the library exp: the fp64 pipe is what bounds the covariance assembly (30 exponentials
// per distance), the categorical kernel (21 per site) and the GRU gates.
__device__ const double DRP_EXP2_TAB[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};

__device__ __forceinline__ double drp_exp(double x)
{
    x = fmin(fmax(x, -700.0), 700.0);
    // t = rint(x * 32/ln2) and n = (int)t through the 1.5 * 2^52 shift: two fp64-pipe operations instead of FRND.F64 +
    // F2I.F64, which run on the XU pipe at a small fraction of the fp64 rate (the GRU gate phase was XU-bound on them)
    const double sh = fma(x, 46.166241308446828384, 6755399441055744.0);
    const int n = __double2loint(sh);
    const double t = sh - 6755399441055744.0;
    double r = fma(t, -0.021660849219188094, x);               // ln2/32, leading 27 bits (t * hi is exact)
    r = fma(t, -1.733101967801894e-10, r);                     // ln2/32, remainder
    double q = 1.3888888888888889e-03;                         // 1/720
    q = fma(q, r, 8.3333333333333332e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double T = __ldg(DRP_EXP2_TAB + (n & 31));
    const double v = fma(T, q * r, T);
    return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
}
