// Microbenchmark: throughput of tcgen05.ld (tensor memory -> registers) for the shapes an epilogue can use, with 16
// epilogue warps reading concurrently (as oz_syrk_kernel does) and optionally a stream of 128x64x32 int8 MMAs running
// on the tensor pipe at the same time.  Question: the tcgen05 update pays ~12 k cycles per 128x64 tile that do not depend
// on the contraction length - is that the 12 x tcgen05.ld.16x256b.x2 (+ wait) each epilogue warp issues per tile?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_rate tmem_ld_rate.cu && ./tmem_ld_rate
//
// Reported: cycles for ONE "tile" = every warp reads its 32 rows x 16 columns of S = 6 accumulators of 64 columns
// (192 KB of tensor memory per CTA), for each load shape, alone and under concurrent MMAs.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_i8(int N, int M) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 lanes x 16 columns (accumulator-fragment layout), 8 registers
__device__ __forceinline__ void ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
// 32 lanes x 16 columns (thread = lane/row), 16 registers
__device__ __forceinline__ void ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}

constexpr int EPI_WARPS = 16;
constexpr int REPS = 64;         // tiles per measurement

// mode 0: 16x256b.x2 + wait after every load (what oz_syrk_kernel does)   mode 1: both halves, then one wait
// mode 2: 32x32b.x16 + wait (one load per accumulator)                     mode 3: all 6 accumulators (32x32b), one wait
template <int MODE>
__device__ uint32_t read_tiles(uint32_t tmem, int q, int h, int reps)
{
    uint32_t sink = 0;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * h);
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll
            for (int d = 0; d < 6; ++d)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t a[8];
                    ld_16x256b_x2(taddr + ((uint32_t)(16 * half) << 16) + (uint32_t)(d * 64), a);
                    ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) sink ^= a[i];
                }
        } else if (MODE == 1) {
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                uint32_t a[8], b[8];
                ld_16x256b_x2(taddr + (uint32_t)(d * 64), a);
                ld_16x256b_x2(taddr + (16u << 16) + (uint32_t)(d * 64), b);
                ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) sink ^= a[i] ^ b[i];
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                uint32_t a[16];
                ld_32x32b_x16(taddr + (uint32_t)(d * 64), a);
                ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) sink ^= a[i];
            }
        } else {
            uint32_t a[6][16];
#pragma unroll
            for (int d = 0; d < 6; ++d) ld_32x32b_x16(taddr + (uint32_t)(d * 64), a[d]);
            ld_wait();
#pragma unroll
            for (int d = 0; d < 6; ++d)
#pragma unroll
                for (int i = 0; i < 16; ++i) sink ^= a[d][i];
        }
    }
    return sink;
}

__global__ void __launch_bounds__((EPI_WARPS + 1) * 32, 1) ld_kernel(long long* out, uint32_t* sinkbuf)
{
    extern __shared__ uint8_t raw_smem[];
    const uint32_t raw = smem_u32(raw_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = raw_smem + (base - raw);
    const uint32_t sA = base, sB = base + 16384, bar = sB + 8192;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + 16384 + 8192 + 16);
    volatile int* flag = reinterpret_cast<volatile int*>(gen + 16384 + 8192 + 32);      // [0] stop MMAs, [1] MMAs issued
    for (int i = threadIdx.x; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        flag[0] = 0;
        flag[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    uint32_t sink = 0;
    uint32_t phase = 0;              // parity of the commit barrier: persists across the measurements
    // 8 measurements: modes 0..3 without MMAs, then modes 0..3 with the MMA warp issuing into columns 384..447
    for (int m = 0; m < 8; ++m) {
        const bool with_mma = m >= 4;
        __syncthreads();
        if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
        __syncthreads();
        const long long t0 = clock64();
        if (warp < EPI_WARPS) {
            const int q = warp & 3, h = warp >> 2;
            switch (m & 3) {
                case 0: sink ^= read_tiles<0>(tmem, q, h, REPS); break;
                case 1: sink ^= read_tiles<1>(tmem, q, h, REPS); break;
                case 2: sink ^= read_tiles<2>(tmem, q, h, REPS); break;
                default: sink ^= read_tiles<3>(tmem, q, h, REPS); break;
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0 && threadIdx.x == 0) out[m] = t1 - t0;
            __syncwarp();
            if (lane == 0) atomicAdd((int*)&flag[0], 1);           // 16 = every epilogue warp is done
        } else if (with_mma) {
            // MMA warp: 128x64x32 int8 SS MMAs into columns 384.. until the epilogue warps are done (bounded)
            const uint32_t idesc = idesc_i8(64, 128);
            int issued = 0;
            for (int it = 0; it < 20000; ++it) {
                if (elect_one()) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) mma_i8_ss(tmem + 384 + 64 * (k4 & 1), umma_desc(sA) + 2 * k4, umma_desc(sB) + 2 * k4, idesc, 1);
                }
                __syncwarp();
                issued += 4;
                if ((it & 15) == 15) {                             // keep the queue bounded: wait for completion every 64 MMAs
                    if (elect_one()) commit(bar);
                    __syncwarp();
                    long long spins = 0;
                    while (!mbar_try_wait(bar, phase)) if (++spins > 2000000ll) break;
                    phase ^= 1;
                    if (flag[0] >= EPI_WARPS || spins > 2000000ll) break;      // done, or a lost commit: give up
                }
            }
            if (elect_one()) commit(bar);
            __syncwarp();
            long long spins = 0;
            while (!mbar_try_wait(bar, phase)) if (++spins > 2000000ll) break;
            phase ^= 1;
            const long long t1 = clock64();
            if (blockIdx.x == 0 && lane == 0) {
                out[8 + m] = issued;
                out[16 + m] = t1 - t0;
            }
        }
    }
    if (sink == 0x12345678u) sinkbuf[threadIdx.x] = sink;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv)
{
    const int ctas = argc > 1 ? atoi(argv[1]) : 1;
    long long* dout;
    uint32_t* dsink;
    cudaMalloc(&dout, 32 * sizeof(long long));
    cudaMemset(dout, 0, 32 * sizeof(long long));
    cudaMalloc(&dsink, 4096);
    const int smem = 1024 + 16384 + 8192 + 64;
    cudaFuncSetAttribute(ld_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ld_kernel<<<ctas, (EPI_WARPS + 1) * 32, smem>>>(dout, dsink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    long long r[32];
    cudaMemcpy(r, dout, sizeof(r), cudaMemcpyDeviceToHost);
    const char* names[] = {"16x256b.x2 + wait each (current epilogue)", "16x256b.x2 both halves, one wait      ",
                           "32x32b.x16 + wait per accumulator      ", "32x32b.x16 x 6 accumulators, one wait  "};
    printf("# %d CTA(s), 16 epilogue warps; cycles per 128x64 tile of 6 int32 accumulators (192 KB of tensor memory)\n", ctas);
    for (int m = 0; m < 8; ++m) {
        printf("%s  %s  %8.0f cycles/tile", names[m & 3], m >= 4 ? "with MMAs" : "alone    ", (double)r[m] / REPS);
        if (m >= 4) printf("   (%lld MMAs of 128x64x32 in %lld cycles = %.1f cycles/MMA)", r[8 + m], r[16 + m], (double)r[16 + m] / (double)r[8 + m]);
        printf("\n");
    }
    return 0;
}
