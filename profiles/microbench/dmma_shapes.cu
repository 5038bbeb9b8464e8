// Microbenchmark: fp64 mma.sync shapes on B200 - throughput of m8n8k4 vs m16n8k16 and a layout check of m16n8k16.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_shapes dmma_shapes.cu && ./dmma_shapes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}

// ---- layout probe: A[16x16] (row-major), B[16x8] as B[k][n]; every lane reports which elements its registers hold
__global__ void probe16816(const double* A, const double* B, double* C, int* amap, int* bmap)
{
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    // hypothesis: a_i: row = gid + 8*(i&1), col = tig + 4*(i>>1);  b_i: k = tig + 4*i, n = gid;  c: rows gid / gid+8, cols 2tig,+1
    double a[8], b[4], c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) a[i] = A[(gid + 8 * (i & 1)) * 16 + tig + 4 * (i >> 1)];
    for (int i = 0; i < 4; ++i) b[i] = B[(tig + 4 * i) * 8 + gid];
    dmma16816(c, a, b);
    C[gid * 8 + 2 * tig] = c[0];
    C[gid * 8 + 2 * tig + 1] = c[1];
    C[(gid + 8) * 8 + 2 * tig] = c[2];
    C[(gid + 8) * 8 + 2 * tig + 1] = c[3];
    (void)amap; (void)bmap;
}
__global__ void probe1684(const double* A, const double* B, double* C)
{
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    double a[2], c[4] = {0, 0, 0, 0};
    a[0] = A[gid * 4 + tig];
    a[1] = A[(gid + 8) * 4 + tig];
    const double b = B[tig * 8 + gid];
    dmma1684(c, a, b);
    C[gid * 8 + 2 * tig] = c[0];
    C[gid * 8 + 2 * tig + 1] = c[1];
    C[(gid + 8) * 8 + 2 * tig] = c[2];
    C[(gid + 8) * 8 + 2 * tig + 1] = c[3];
}

template <int MODE>
__global__ void __launch_bounds__(256) tput(double* out, int iters)
{
    double acc[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = 0.5 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { dmma884(acc[i][0], acc[i][1], a[i], b[i & 3]); dmma884(acc[i][2], acc[i][3], a[i], b[(i + 1) & 3]); }
            if (MODE == 1) dmma16816(acc[i], a, b);
            if (MODE == 2) { double a2[2] = {a[i], a[(i + 1) & 7]}; dmma1684(acc[i], a2, b[i & 3]); }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- legacy int8 mma.sync (IMMA) throughput, alone and next to fp64 FMAs (does it run beside the fp64 pipe?)
__device__ __forceinline__ void imma16832(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE>   // 0: all warps IMMA, 2: warps 0-3 IMMA + warps 4-7 DFMA
__global__ void __launch_bounds__(256) imix(double* out, int iters)
{
    const int warp = threadIdx.x >> 5;
    const bool do_imma = MODE == 0 || (warp >> 2) == 0;
    int acc[8][4];
    double f[16];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int i = 0; i < 16; ++i) f[i] = 1.0 + i * 1e-3;
    unsigned a[4] = {0x01020304u + threadIdx.x, 0x05060708u, 0x090a0b0cu, 0x0d0e0f10u}, b[2] = {0x01010101u, 0x02020202u};
    const double fa = 1.0 + threadIdx.x * 1e-9, fb = 0.999999;
    if (do_imma) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) imma16832(acc[i], a, b);
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], fb, fa);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- do DMMA and DFMA share execution resources?  warps 0-3 run DMMAs, warps 4-7 DFMAs (every SM sub-partition hosts one warp of each kind)
template <int MODE>   // 0: all warps DMMA, 1: all warps DFMA, 2: warps 0-3 DMMA + warps 4-7 DFMA
__global__ void __launch_bounds__(256) mix(double* out, int iters)
{
    const int warp = threadIdx.x >> 5;
    const bool do_dmma = MODE == 0 || (MODE == 2 && (warp >> 2) == 0);   // warps w and w+4 share a sub-partition: one of each kind
    double acc[8][2], f[16];
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int i = 0; i < 16; ++i) f[i] = 1.0 + i * 1e-3;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999;
    if (do_dmma) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], b, a);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double hA[256], hB[128], hC[128], ref[128];
    for (int i = 0; i < 256; ++i) hA[i] = (double)((i * 37) % 101) / 7.0 - 3.0;
    for (int i = 0; i < 128; ++i) hB[i] = (double)((i * 53) % 89) / 5.0 - 4.0;
    double *dA, *dB, *dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    probe16816<<<1, 32>>>(dA, dB, dC, nullptr, nullptr);
    cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += hA[m * 16 + k] * hB[k * 8 + n]; ref[m * 8 + n] = s; double e = s - hC[m * 8 + n]; if (e < 0) e = -e; if (e > err) err = e; }
    printf("m16n8k16 layout check: max err %.3e (%s)\n", err, err < 1e-9 ? "OK" : "MISMATCH");
    probe1684<<<1, 32>>>(dA, dB, dC);   // A read as [16x4], B as [4x8]
    cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    err = 0;
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < 4; ++k) s += hA[m * 4 + k] * hB[k * 8 + n]; double e = s - hC[m * 8 + n]; if (e < 0) e = -e; if (e > err) err = e; }
    printf("m16n8k4  layout check: max err %.3e (%s)\n", err, err < 1e-9 ? "OK" : "MISMATCH");

    int dev = 0, sms = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) tput<0><<<sms * 4, 256>>>(out, iters);
            if (mode == 1) tput<1><<<sms * 4, 256>>>(out, iters);
            if (mode == 2) tput<2><<<sms * 4, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // flops per warp-iteration: mode0: 16 x (8*8*4*2) ; mode1: 8 x (16*8*16*2) ; mode2: 8 x (16*8*4*2)
        const double per = mode == 0 ? 16.0 * 512 : (mode == 1 ? 8.0 * 4096 : 8.0 * 1024);
        const double flops = per * iters * (double)sms * 4 * 8;
        printf("%s: %.3f ms, %.2f TFLOP/s\n", mode == 0 ? "m8n8k4  " : (mode == 1 ? "m16n8k16" : "m16n8k4 "), ms, flops / ms / 1e9);
    }
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) mix<0><<<sms * 4, 256>>>(out, iters);
            if (mode == 1) mix<1><<<sms * 4, 256>>>(out, iters);
            if (mode == 2) mix<2><<<sms * 4, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // per warp-iteration: DMMA warps 8 x 512 flop, DFMA warps 16 x 64 flop
        const double warps = (double)sms * 4 * 8;
        const double dmma_fl = (mode == 1 ? 0.0 : (mode == 0 ? 1.0 : 0.5)) * warps * iters * 8.0 * 512;
        const double dfma_fl = (mode == 0 ? 0.0 : (mode == 1 ? 1.0 : 0.5)) * warps * iters * 16.0 * 64;
        printf("mix mode %d (%s): %.3f ms, DMMA %.2f TFLOP/s, DFMA %.2f TFLOP/s\n", mode,
               mode == 0 ? "all DMMA" : (mode == 1 ? "all DFMA" : "half DMMA / half DFMA"), ms, dmma_fl / ms / 1e9, dfma_fl / ms / 1e9);
    }
    for (int mode = 0; mode < 3; mode += 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) imix<0><<<sms * 4, 256>>>(out, iters);
            if (mode == 2) imix<2><<<sms * 4, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)sms * 4 * 8;
        const double imma_ops = (mode == 0 ? 1.0 : 0.5) * warps * iters * 8.0 * (16 * 8 * 32 * 2);
        const double dfma_fl = (mode == 0 ? 0.0 : 0.5) * warps * iters * 16.0 * 64;
        printf("imma mode %d (%s): %.3f ms, IMMA m16n8k32 %.1f TOP/s (%.0f MAC/clk/SM at 1.965 GHz), DFMA %.2f TFLOP/s\n", mode,
               mode == 0 ? "all IMMA" : "half IMMA / half DFMA", ms, imma_ops / ms / 1e9,
               imma_ops / 2 / (ms * 1e-3) / sms / 1.965e9, dfma_fl / ms / 1e9);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
