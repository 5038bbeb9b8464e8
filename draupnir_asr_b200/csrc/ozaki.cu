// fp64-accurate trailing update of the blocked Cholesky on the INT8 tcgen05 tensor pipe (sm_100a).
//
// The trailing update  C[r][c] -= sum_k P[r][k] P[c][k]  (factor.cu) is the only dense contraction on the OU-prior
// path (torch.linalg.cholesky / cholesky_backward / linalg.inv of the reference: S/models.py:118,167-193).  tcgen05
// has no f64 kind, so the fp64 panel P is split, row by row, into S byte slices of a fixed-point mantissa
//
//     P[r][k] ~= 2^(e_r - 7) * ( q0 + q1 2^-8 + ... + q_{S-1} 2^-8(S-1) ),   q0 in [-128,127] (s8), q_t in [0,255] (u8)
//
// (e_r = exponent of the row maximum, mantissa rounded to nearest at 8S-1 bits), and the products of slices are
// accumulated EXACTLY in int32 by tcgen05.mma.kind::i8 into S TMEM accumulators, one per anti-diagonal d = t + u:
//
//     sum_k P[r][k] P[c][k] ~= 2^(e_r-7) 2^(e_c-7) * sum_{d<S} 2^-8d * ( sum_{t+u=d} sum_k q_t[r][k] q_u[c][k] )
//
// The epilogue recombines the S integers in fp64 (one fma per term, in the order the accumulators complete) and
// applies the update to C.  The only
// error is the mantissa truncation (2^-(8S-1) of the row maximum) and the dropped anti-diagonals d >= S
// (~ S 2^-8S of the product of row maxima): S = 6 gives ~1e-13, S = 7 ~1e-15 relative to |P_r|max |P_c|max, i.e. a
// backward error at fp64 level for the factorisation (DESIGN.md "split-integer trailing update").
//
// Data movement: oz_slice_kernel writes the slices in global memory already in the 128-byte-swizzled K-major shared
// memory image that tcgen05 reads, so every operand tile is ONE contiguous 16 KB / 8 KB block moved by a TMA bulk
// copy (cp.async.bulk, completes on an mbarrier).  oz_syrk_kernel is persistent and warp-specialised:
//   warp 0   TMA producer: the S*KBN A tiles of a 128-row strip stay resident, B tiles stream through a ring
//   warp 1   MMA issuer (one lane): 128x64x32 int8 MMAs into TMEM, tcgen05.commit -> mbarriers
//   warps 2-17 epilogue: prefetch of the C tile, tcgen05.ld of the S accumulators, fp64 recombination, masked write-back
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"
#include "slice.cuh"
#include <stdlib.h>

namespace {

constexpr int OZ_TM = 128;                       // tile rows  (UMMA M)
constexpr int OZ_TN = 64;                        // tile cols  (UMMA N)
constexpr int OZ_ROWB = 128;                     // bytes (= int8 K elements) per swizzled row of a k-block
constexpr int OZ_A_TILE = OZ_TM * OZ_ROWB;       // 16 KB
constexpr int OZ_B_TILE = OZ_TN * OZ_ROWB;       // 8 KB
constexpr int OZ_CHUNK = 8;                      // column tiles per work unit (A strip reloaded per unit)
constexpr int OZ_EPI_WARPS = 16;                 // 4 TMEM lane quadrants x 4 column quarters of the 128x64 tile
constexpr int OZ_THREADS = (2 + OZ_EPI_WARPS) * 32;
constexpr int OZ_SMEM_MAX = 232448;              // 227 KB
constexpr int OZ_MAXS = 8;                       // upper bound on the slice count (barrier storage)
constexpr long long OZ_TIMEOUT_CYCLES = 6000000000LL;   // a stuck barrier traps instead of hanging the device

template <int S, int KBN>
struct OzCfg {
    static constexpr int NA = S * KBN;                                   // resident A tiles per strip
    static constexpr int A_BYTES = NA * OZ_A_TILE;
    static constexpr int NST_RAW = (OZ_SMEM_MAX - 1024 - 320 - A_BYTES) / OZ_B_TILE;
    static constexpr int NST = NST_RAW > 8 ? 8 : NST_RAW;                // B ring stages
    static constexpr int SMEM = 1024 + A_BYTES + NST * OZ_B_TILE + 320;
    static constexpr int TMEM_COLS = (S * OZ_TN <= 64) ? 64 : (S * OZ_TN <= 128) ? 128 : (S * OZ_TN <= 256) ? 256 : 512;
    static_assert(NST >= 2, "A strip does not leave room for the B ring");
    static_assert(S * OZ_TN <= 512, "too many accumulators for TMEM");
};

struct OzParams {
    double* M;                 // fp64 matrices, row-major
    int64_t ld, bs;
    const uint8_t* sl;         // slices [z][S][KBN][Rpad][128 B swizzled]
    int64_t sl_bs;             // bytes per matrix
    const int32_t* scl;        // [z][Rpad]  binary exponent e_r of the row maximum (row scale 2^(e_r - 7))
    int Rpad, rb, c1, R, cmax;
    int nTi, nTj, units_per_z, total_units;
    int slice_c1;              // first live row of the panel (rows of the slice buffer above it are zeros)
    int ti0, tj0;              // first 128-row strip / first 64-column tile of this launch (a column-range part of an update)
    int K;                     // real contraction length (columns of the panel); the k-blocks are zero-padded
    long long* trace;          // event trace of CTA 0 (DRP_DEBUG_HOOKS builds only)
    long long* prof;           // per-CTA wait-time sums [grid][16] (DRP_DEBUG_HOOKS builds only)
    int* counter;              // next work unit of THIS launch (both counters are zeroed by the slicing kernel of the update)
    int counter_idx;           // 0 / 1: which of the two
    const int* zmap;           // precision gate (factor.cuh): [0] matrices on this path, [2..] their indices; null = all z
};

// ------------------------------------------------ PTX wrappers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 0xFFF) == 0 && clock64() - t0 > OZ_TIMEOUT_CYCLES) __trap();
    }
}
// mbar_wait that adds the cycles it blocked to `acc` (phase accounting of the roles; two clock reads per wait)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc)
{
#ifdef DRP_DEBUG_HOOKS
    const long long t = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t;
#else
    (void)acc;
    mbar_wait(bar, parity);
#endif
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// one lane of the (fully converged) warp
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by ONE thread for the CTA (SASS: UTCIMMA)
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 lanes x 16 columns in the accumulator-fragment layout: thread t gets rows (t/4, t/4+8) of the 16-lane window and,
// for each 8-column block k, columns 8k + 2(t%4), +1:  v[4k+0..1] = row t/4, v[4k+2..3] = row t/4+8.
// A warp-wide 16-byte access to those elements touches 8 rows x 64 contiguous bytes (full sectors).
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: 8-row groups 1024 B apart (SBO), version 1 (sm_100), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::i8 instruction descriptor: D = s32, A/B signed (1) or unsigned (0), both K-major, N = 64, M = 128
__device__ __forceinline__ uint32_t umma_idesc(uint32_t a_signed, uint32_t b_signed)
{
    return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(OZ_TN >> 3) << 17) | ((uint32_t)(OZ_TM >> 4) << 24);
}

// 2^ex as a double (exponent-field construction; far out-of-range products of row scales go through scalbn)
__device__ __forceinline__ double oz_pow2(int ex)
{
    return (ex > -1000 && ex < 1000) ? __hiloint2double((ex + 1023) << 20, 0) : scalbn(1.0, ex);
}

// exact int32 -> fp64 without the conversion pipe: 2^52 + (a + 2^31) assembled in the mantissa, then one subtraction
__device__ __forceinline__ double i32_to_f64(uint32_t a)
{
    return __hiloint2double(0x43300000, (int)(a ^ 0x80000000u)) - 4503601774854144.0;
}

// debug trace: (event id, clock) pairs appended by single threads of CTA 0
#ifndef DRP_DEBUG_HOOKS
#define OZ_TRACE(id) do { } while (0)
#else
#define OZ_TRACE(id)                                                                   \
    do {                                                                               \
        if (p.trace && blockIdx.x == 0) {                                              \
            const unsigned long long _k = atomicAdd((unsigned long long*)p.trace, 1ull); \
            if (_k < 4000) {                                                           \
                p.trace[2 + 2 * _k] = (id);                                            \
                p.trace[3 + 2 * _k] = clock64();                                       \
            }                                                                          \
        }                                                                              \
    } while (0)
#endif

// work unit u -> (matrix, 128-row strip, first column tile, number of column tiles); strips longest first
__device__ __forceinline__ void oz_decode(const OzParams& p, int u, int& zz, int& ti, int& tj0, int& ntj)
{
    zz = u / p.units_per_z;
    int v = u - zz * p.units_per_z;
    if (p.zmap) zz = p.zmap[2 + zz];
    tj0 = p.tj0;
    ntj = 0;
    for (ti = p.nTi - 1; ti >= p.ti0; --ti) {
        const int tot = min(2 * ti + 2, p.nTj) - p.tj0;        // column tiles [p.tj0, min(2 ti + 2, nTj)) of this strip
        const int nch = tot > 0 ? (tot + OZ_CHUNK - 1) / OZ_CHUNK : 0;
        if (v < nch) {
            tj0 = p.tj0 + v * OZ_CHUNK;
            ntj = min(OZ_CHUNK, tot - v * OZ_CHUNK);
            return;
        }
        v -= nch;
    }
    ti = p.ti0;
}

// ---------------------------------------------------------------------------------------------------------------
// Slicing: one warp per row of the panel P = M[rb .. rb+Rpad) x [k0, k0+K).  Rows outside [c1, R) become zeros.
// ---------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ Mb, int64_t ld, int64_t bs, int k0, int Kcols, int KBN,
                                                       int rb, int c1, int R, int Rpad, uint8_t* __restrict__ sl,
                                                       int64_t sl_bs, int32_t* __restrict__ scl, int* __restrict__ counter,
                                                       const int* __restrict__ zmap)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) counter[0] = counter[1] = 0;   // work-unit counters of the (up to two) update launches that follow
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = blockIdx.x * 8 + warp;
    int zz = blockIdx.y;
    if (zmap) {                                    // matrices the precision gate keeps on the fp64 path are not sliced
        if (zz >= zmap[0]) return;
        zz = zmap[2 + zz];
    }
    if (rr >= Rpad) return;
    const int r = rb + rr;
    const bool live = (r >= c1 && r < R);
    const double* src = Mb + (int64_t)zz * bs + (int64_t)r * ld + k0 + 4 * lane;
    double x[8];
    double m = 0.0;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double v = (live && kb * 128 + 4 * lane + j < Kcols) ? src[kb * 128 + j] : 0.0;
            x[kb * 4 + j] = v;
            m = fmax(m, fabs(v));
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int e = 0;
    if (m > 0.0 && m < 1e300) frexp(m, &e);
    if (lane == 0) scl[(int64_t)zz * Rpad + rr] = e;
    uint8_t* dst = sl + (int64_t)zz * sl_bs;
    const int chunk = ((lane >> 2) ^ (rr & 7)) << 4;       // 16-byte chunk of the row after the 128B swizzle
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
        if (kb >= KBN) break;
        long long X[4];
        const double xk[4] = {x[kb * 4], x[kb * 4 + 1], x[kb * 4 + 2], x[kb * 4 + 3]};
        drp_quant4(xk, m, e, S, X);
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int64_t off = ((int64_t)(t * KBN + kb) * Rpad + rr) * OZ_ROWB + chunk + ((4 * lane) & 15);
            *reinterpret_cast<uint32_t*>(dst + off) = drp_pick4(X, S - 1 - t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The update kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int S, int KBN>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_syrk_kernel(const OzParams p)
{
    using Cfg = OzCfg<S, KBN>;
    constexpr int NST = Cfg::NST;
    extern __shared__ uint8_t oz_smem_raw[];
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;              // swizzle atoms need 1024-byte alignment
    uint8_t* gen = oz_smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + Cfg::A_BYTES;
    const uint32_t bars = sB + NST * OZ_B_TILE;
    // mbarriers: A strip full/empty, accumulators drained, accumulator d complete (one per anti-diagonal), B ring
    const uint32_t bar_a_full = bars, bar_a_empty = bars + 8, bar_acc_empty = bars + 16, bar_acc_full = bars + 24;
    const uint32_t bar_b_full = bars + 24 + 8 * OZ_MAXS, bar_b_empty = bar_b_full + 8 * NST;
    // work units are handed out dynamically (atomic counter): the producer publishes each unit id in a 4-slot queue
    // (bar_unit[slot] = "slot written"); the pipeline depth keeps every role within 2 units of the producer
    const uint32_t bar_unit = bar_b_empty + 8 * NST;
    volatile int* unit_q = reinterpret_cast<volatile int*>(gen + Cfg::A_BYTES + NST * OZ_B_TILE + 24 + 8 * OZ_MAXS + 16 * NST + 32);
    volatile uint32_t* tmem_slot =
        reinterpret_cast<volatile uint32_t*>(gen + Cfg::A_BYTES + NST * OZ_B_TILE + 24 + 8 * OZ_MAXS + 16 * NST + 48);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int d = 0; d < S; ++d) mbar_init(bar_acc_full + 8 * d, 1);
        mbar_init(bar_acc_empty, OZ_EPI_WARPS);
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, 1);
        }
        for (int j = 0; j < 4; ++j) mbar_init(bar_unit + 8 * j, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            uint32_t a_phase = 0, b_phase = 0;
            int st = 0;
            long long w_a = 0, w_b = 0;
            const long long t_start = clock64();
            const int total_units = p.zmap ? p.zmap[0] * p.units_per_z : p.total_units;
            for (int i = 0;; ++i) {
                int u = atomicAdd(p.counter, 1);
                if (u >= total_units) u = -1;
                unit_q[i & 3] = u;
                mbar_arrive(bar_unit + 8 * (i & 3));
                if (u < 0) break;
                int zz, ti, tj0, ntj;
                oz_decode(p, u, zz, ti, tj0, ntj);
                const uint8_t* src = p.sl + (int64_t)zz * p.sl_bs;
                mbar_wait_t(bar_a_empty, a_phase ^ 1, w_a);
                OZ_TRACE(100);
                mbar_expect_tx(bar_a_full, Cfg::A_BYTES);
                for (int a = 0; a < Cfg::NA; ++a)
                    bulk_g2s(sA + a * OZ_A_TILE, src + ((int64_t)a * p.Rpad + (int64_t)ti * OZ_TM) * OZ_ROWB, OZ_A_TILE,
                             bar_a_full);
                a_phase ^= 1;
                for (int t = 0; t < ntj; ++t) {
                    const int64_t cc0 = (int64_t)(tj0 + t) * OZ_TN;
                    for (int b = 0; b < Cfg::NA; ++b) {
                        mbar_wait_t(bar_b_empty + 8 * st, b_phase ^ 1, w_b);
                        OZ_TRACE(200 + b);
                        mbar_expect_tx(bar_b_full + 8 * st, OZ_B_TILE);
                        bulk_g2s(sB + st * OZ_B_TILE, src + ((int64_t)b * p.Rpad + cc0) * OZ_ROWB, OZ_B_TILE,
                                 bar_b_full + 8 * st);
                        if (++st == NST) {
                            st = 0;
                            b_phase ^= 1;
                        }
                    }
                }
            }
#ifdef DRP_DEBUG_HOOKS
            if (p.prof) {
                long long* o = p.prof + 16 * blockIdx.x;
                o[0] = clock64() - t_start;   // producer: total, waiting for the A strip to be released, for a free B stage
                o[1] = w_a;
                o[2] = w_b;
            }
#else
            (void)t_start; (void)w_a; (void)w_b;
#endif
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ==============================
        // The whole warp walks the loop convergently and ONE elected lane issues (elect.sync): tcgen05.mma inside
        // divergent `lane == 0` code is wrapped by the compiler in a per-thread ELECT/branch loop that costs more
        // issue cycles than the 32-cycle MMA itself lasts.
        uint32_t a_phase = 0, b_phase = 0, acc_phase = 0;
        int st = 0;
        const bool full_k = p.K == KBN * OZ_ROWB;      // no ragged last k-block
        long long w_unit = 0, w_a = 0, w_acc = 0, w_b = 0, n_tiles = 0;
        const long long t_start = clock64();
        for (int i = 0;; ++i) {
            mbar_wait_t(bar_unit + 8 * (i & 3), (uint32_t)(i >> 2) & 1u, w_unit);
            const int u = unit_q[i & 3];
            if (u < 0) break;
            int zz, ti, tj0, ntj;
            oz_decode(p, u, zz, ti, tj0, ntj);
            mbar_wait_t(bar_a_full, a_phase, w_a);
            a_phase ^= 1;
            tc_fence_after();
            if (lane == 0) OZ_TRACE(300);
            for (int t = 0; t < ntj; ++t) {
                mbar_wait_t(bar_acc_empty, acc_phase ^ 1, w_acc);
                ++n_tiles;
                tc_fence_after();
                if (lane == 0) OZ_TRACE(310);
#pragma unroll
                for (int ub = 0; ub < S; ++ub) {               // slice of B (columns of the update)
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) {
                        mbar_wait_t(bar_b_full + 8 * st, b_phase, w_b);
                        tc_fence_after();
                        if (lane == 0) OZ_TRACE(400 + ub * KBN + kb);
                        if (elect_one()) {
                            const uint64_t bdesc = umma_desc(sB + st * OZ_B_TILE);
                            const int k4n = full_k ? 4 : min(4, (p.K - OZ_ROWB * kb + 31) >> 5);   // 32-column steps with data
                            // k-step outer, A slice inner: consecutive MMAs accumulate into different TMEM tiles
                            // (anti-diagonals d = ta + ub).  The order does not matter for the rate: 48-51 cycles per
                            // 128x64x32 MMA either way in isolation (profiles/microbench/imma_rate.cu; the "accumulate
                            // latency" an earlier comment quoted was the trace hook's own cost), and the A-slice-outer
                            // order timed 0-2 % slower in the kernel (profiles/r01/oz_syrk_ab_variants.txt).
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                if (!full_k && k4 >= k4n) break;
#pragma unroll
                                for (int ta = 0; ta + ub < S; ++ta) {  // slice of A
                                    const uint64_t adesc = umma_desc(sA + (ta * KBN + kb) * OZ_A_TILE);
                                    const uint32_t idesc = umma_idesc(ta == 0, ub == 0);
                                    const uint32_t dcol = tmem + (uint32_t)((ta + ub) * OZ_TN);
                                    umma_i8(dcol, adesc + 2 * k4, bdesc + 2 * k4, idesc, (ub | kb | k4) != 0);
                                }
                            }
                            umma_commit(bar_b_empty + 8 * st);
                            // anti-diagonal d = ub only receives products with B slices <= ub: complete after its last
                            // k-block, so the epilogue starts on it while the remaining slice pairs are multiplied
                            if (kb == KBN - 1) umma_commit(bar_acc_full + 8 * ub);
                            if (kb == KBN - 1 && ub == S - 1 && t == ntj - 1) umma_commit(bar_a_empty);
                        }
                        __syncwarp();
                        if (++st == NST) {
                            st = 0;
                            b_phase ^= 1;
                        }
                    }
                }
                acc_phase ^= 1;
            }
        }
#ifdef DRP_DEBUG_HOOKS
        if (p.prof && lane == 0) {
            long long* o = p.prof + 16 * blockIdx.x;
            o[3] = clock64() - t_start;       // MMA issuer: total, waiting for a unit, the A strip, drained accumulators, a B block
            o[4] = w_unit;
            o[5] = w_a;
            o[6] = w_acc;
            o[7] = w_b;
            o[8] = n_tiles;
        }
#else
        (void)t_start; (void)w_unit; (void)w_a; (void)w_acc; (void)w_b; (void)n_tiles;
#endif
    } else {
        // ============================ epilogue ================================
        // warp (q, h) owns rows 32q.. and columns 16h.. of the 128x64 tile; thread t holds, for row select i = 0..3,
        // row 32q + t/4 + 8i and, for k = 0..1, the column pair 16h + 8k + 2(t%4), +1.  The C values are prefetched to
        // L2 BEFORE waiting for the accumulators (their DRAM latency hides behind the tile's MMAs) and the
        // read-modify-write happens after the accumulators have been released, overlapping the next tile's MMAs.
        const int q = warp & 3;                    // TMEM lane quadrant this warp may read
        const int h = (warp - 2) >> 2;             // which 16 of the tile's 64 columns
        const int tr = lane >> 2, tc = 2 * (lane & 3);
        const bool vec_ok = ((p.ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.M) & 15) == 0) && ((p.bs & 1) == 0);
        uint32_t acc_phase = 0;
        long long w_unit = 0, w_acc = 0;
        const long long t_start = clock64();
        for (int i = 0;; ++i) {
            mbar_wait_t(bar_unit + 8 * (i & 3), (uint32_t)(i >> 2) & 1u, w_unit);
            const int u = unit_q[i & 3];
            if (u < 0) break;
            int zz, ti, tj0, ntj;
            oz_decode(p, u, zz, ti, tj0, ntj);
            const int32_t* sc = p.scl + (int64_t)zz * p.Rpad;
            const int rr0 = ti * OZ_TM + q * 32 + tr;
            int er[4];                                  // row exponents: one batched load per unit
#pragma unroll
            for (int i = 0; i < 4; ++i) er[i] = sc[rr0 + 8 * i];
            const int r0 = p.rb + rr0;                  // this thread's rows are r0 + 8 i
            double* const cbase = p.M + (int64_t)zz * p.bs + (int64_t)r0 * p.ld;
            const int64_t rstep = 8 * p.ld;
#define OZ_CLIMIT(i) ((r0 + 8 * (i) >= p.c1 && r0 + 8 * (i) < p.R) ? min(r0 + 8 * (i) + 1, p.cmax) : 0)
            // Interior fast path (warp-uniform): the warp's 32 rows x 16 columns lie entirely inside the updated region
            // (most tiles off the diagonal) and every scale exponent is far from the fp64 limits, so the C update needs
            // no per-element masks, no range checks and no divergence.  Per tile that removes ~40 % of the epilogue's
            // instructions (compares, selects, branches), which - not the tensor pipe - is what a tile's time is made of.
            const int rw0 = p.rb + ti * OZ_TM + q * 32;   // first row of this warp's window
            const bool rows_in = vec_ok && rw0 >= p.c1 && rw0 + 31 < p.R;
            const bool er_safe = (abs(er[0]) | abs(er[1]) | abs(er[2]) | abs(er[3])) < 256;
            for (int t = 0; t < ntj; ++t) {
                const int cc0 = (tj0 + t) * OZ_TN + 16 * h + tc;      // local column of this thread's first pair
                int ec[4];                                 // column exponents of this thread's two column pairs
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    ec[2 * k] = sc[cc0 + 8 * k];
                    ec[2 * k + 1] = sc[cc0 + 8 * k + 1];
                }
                const int cw0 = p.rb + (tj0 + t) * OZ_TN + 16 * h;     // first column of this warp's window
                const bool geom = rows_in && cw0 >= p.c1 && cw0 + 15 < min(rw0 + 1, p.cmax);
                // pull this thread's part of the C tile towards L2 while the tile's MMAs run
                if (geom) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) prefetch_l2(cbase + i * rstep + (p.rb + cc0 + 8 * k));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int col = p.rb + cc0 + 8 * k;
                            if (col < OZ_CLIMIT(i)) prefetch_l2(cbase + i * rstep + col);
                        }
                }
                double v[4][4];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * h);
#pragma unroll
                for (int d = 0; d < S; ++d) {                    // in the order the anti-diagonals complete
                    mbar_wait_t(bar_acc_full + 8 * d, acc_phase, w_acc);
                    tc_fence_after();
                    if (warp == 2 && lane == 0) OZ_TRACE(500 + d);
                    const double w = 1.0 / (double)(1ull << (8 * d));                 // 2^-8d
#pragma unroll
                    for (int half = 0; half < 2; ++half) {      // lanes 32q + 16 half ..: rows i = 2 half, 2 half + 1
                        uint32_t a[8];
                        tmem_ld_16x256b_x2(taddr + ((uint32_t)(16 * half) << 16) + (uint32_t)(d * OZ_TN), a);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 2; ++k)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const double x0 = i32_to_f64(a[4 * k + e]), x1 = i32_to_f64(a[4 * k + 2 + e]);
                                if (d == 0) {
                                    v[2 * half][2 * k + e] = x0;
                                    v[2 * half + 1][2 * k + e] = x1;
                                } else {
                                    v[2 * half][2 * k + e] = fma(x0, w, v[2 * half][2 * k + e]);
                                    v[2 * half + 1][2 * k + e] = fma(x1, w, v[2 * half + 1][2 * k + e]);
                                }
                            }
                    }
                }
                acc_phase ^= 1;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty);       // accumulators may be overwritten by the next tile
                if (warp == 2 && lane == 0) OZ_TRACE(600);
                const bool ec_safe = (abs(ec[0]) | abs(ec[1]) | abs(ec[2]) | abs(ec[3])) < 256;
                if (geom && __all_sync(0xffffffffu, er_safe && ec_safe)) {
                    double2 c[4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            c[i][k] = *reinterpret_cast<const double2*>(cbase + i * rstep + (p.rb + cc0 + 8 * k));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            // 2^(er + ec - 14): both exponents are below 256 in magnitude, no range check needed
                            const double s0 = __hiloint2double((er[i] + ec[2 * k] - 14 + 1023) << 20, 0);
                            const double s1 = __hiloint2double((er[i] + ec[2 * k + 1] - 14 + 1023) << 20, 0);
                            *reinterpret_cast<double2*>(cbase + i * rstep + (p.rb + cc0 + 8 * k)) =
                                make_double2(fma(-v[i][2 * k], s0, c[i][k].x), fma(-v[i][2 * k + 1], s1, c[i][k].y));
                        }
                } else {
                    double2 c[4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int col = p.rb + cc0 + 8 * k;
                            const int cl = OZ_CLIMIT(i);
                            const double* crow = cbase + i * rstep;
                            c[i][k] = make_double2(0.0, 0.0);
                            if (vec_ok && col >= p.c1 && col + 1 < cl) {
                                c[i][k] = *reinterpret_cast<const double2*>(crow + col);
                            } else {
                                if (col >= p.c1 && col < cl) c[i][k].x = crow[col];
                                if (col + 1 >= p.c1 && col + 1 < cl) c[i][k].y = crow[col + 1];
                            }
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int col = p.rb + cc0 + 8 * k;
                            const int cl = OZ_CLIMIT(i);
                            double* crow = cbase + i * rstep;
                            const double w0 = c[i][k].x - v[i][2 * k] * oz_pow2(er[i] + ec[2 * k] - 14);
                            const double w1 = c[i][k].y - v[i][2 * k + 1] * oz_pow2(er[i] + ec[2 * k + 1] - 14);
                            if (vec_ok && col >= p.c1 && col + 1 < cl) {
                                *reinterpret_cast<double2*>(crow + col) = make_double2(w0, w1);
                            } else {
                                if (col >= p.c1 && col < cl) crow[col] = w0;
                                if (col + 1 >= p.c1 && col + 1 < cl) crow[col + 1] = w1;
                            }
                        }
                }
            }
        }
#ifdef DRP_DEBUG_HOOKS
        if (p.prof && warp == 2 && lane == 0) {
            long long* o = p.prof + 16 * blockIdx.x;
            o[9] = clock64() - t_start;       // epilogue warp 2: total, waiting for a unit, for completed accumulators
            o[10] = w_unit;
            o[11] = w_acc;
        }
#else
        (void)t_start; (void)w_unit; (void)w_acc;
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// Narrowest update (columns) that takes the tensor path.  DRP_OZ_MIN_COLS overrides.
static int oz_min_cols()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_OZ_MIN_COLS");
        v = e ? atoi(e) : 64;
        if (v < 64) v = 64;
    }
    return v;
}

static bool oz_yield()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_LOOKAHEAD_YIELD");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

template <int S, int KBN>
int oz_launch(const OzParams& p, int z, int k0, int Kcols, cudaStream_t st, double flops, bool do_slice)
{
    using Cfg = OzCfg<S, KBN>;
    static int n_sm = 0;
    if (drp_first_on_device(8 + S * 2 + KBN)) {
        cudaFuncSetAttribute(oz_syrk_kernel<S, KBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    if (do_slice) {
        DRP_LAUNCH(DRP_K_SLICE, (double)z * p.Rpad * KBN * 128 * (8.0 + S), st);
        oz_slice_kernel<S><<<dim3((p.Rpad + 7) / 8, z), 256, 0, st>>>(p.M, p.ld, p.bs, k0, Kcols, KBN, p.rb, p.slice_c1, p.R, p.Rpad,
                                                                     const_cast<uint8_t*>(p.sl), p.sl_bs,
                                                                     const_cast<int32_t*>(p.scl), p.counter - p.counter_idx, p.zmap);
    }
    {
        DRP_LAUNCH(DRP_K_SYRK_TC, flops, st);
        // counter_idx 1 = the look-ahead part of an update, running beside the next panel's factorisation chain: one CTA per
        // work unit instead of one persistent CTA per SM, so that SMs are handed back every ~50 us and the (higher-priority)
        // chain kernels get them; CTAs that find the counter exhausted exit at once
        const int grid = (p.counter_idx == 1 && oz_yield()) ? p.total_units : (p.total_units < n_sm ? p.total_units : n_sm);
        oz_syrk_kernel<S, KBN><<<grid, OZ_THREADS, Cfg::SMEM, st>>>(p);
    }
    return 0;
}

}  // namespace

// The phase-accounting / event-trace hooks of the update kernel are process-global debug state and therefore NOT part of
// the product library: they exist only in a -DDRP_DEBUG_HOOKS build (profiles/build_debug.sh -> libdraupnir_b200_debug.so),
// which the profiling scripts under profiles/ load instead.
#ifdef DRP_DEBUG_HOOKS
static long long* g_oz_trace = nullptr;
static long long* g_oz_prof = nullptr;
// device buffer of >= 16 * (number of SMs) int64; every CTA of the following updates stores the cycles its
// producer / MMA issuer / epilogue warp 2 spent in total and blocked on each barrier (layout: see the kernel).
DRP_API void drp_ozaki_set_prof(void* buf) { g_oz_prof = (long long*)buf; }
// device buffer of >= 8002 int64 ([0] = event count) that CTA 0 of every following update appends to.
DRP_API void drp_ozaki_set_trace(void* buf) { g_oz_trace = (long long*)buf; }
#else
static long long* const g_oz_trace = nullptr;
static long long* const g_oz_prof = nullptr;
#endif

int64_t drp_zmap_bytes(int z) { return (((int64_t)z * 8 + 8 + 255) / 256) * 256; }

// Bytes of ONE slice buffer (slices, row scales, the two work-unit counters) for matrices with at most R_full rows.
static int64_t oz_buffer_bytes(int R_full, int z)
{
    const int64_t Rpad = (((int64_t)R_full + 127) / 128 + 1) * 128;
    return (((int64_t)z * Rpad * (12 * OZ_ROWB + 8) + 2048) + 1023) / 1024 * 1024;
}

// Bytes of workspace the tensor path may need for matrices with at most R_full rows: TWO slice buffers (buffer 0: the
// updates that leave an outer panel, buffer 1: the updates inside a panel - with look-ahead the latter are sliced while the
// former are still being multiplied) and the precision gate's index lists at the very end.
int64_t drp_ozaki_workspace_bytes(int R_full, int z)
{
    return 2 * oz_buffer_bytes(R_full, z) + drp_zmap_bytes(z) + drp_linv_bytes(z);
}

// Inverses of the (up to four) 64 x 64 diagonal blocks of the panel being factored, per matrix (factor.cu: fused panel
// kernels); the last region of the workspace.
int64_t drp_linv_bytes(int z) { return (int64_t)z * 4 * DRP_NB * DRP_NB * 8 + 256; }

// Largest condition-number bound a matrix may have to take the split-integer update: eps_S * cond <= 1e-5 with the
// backward error eps_S ~ 2^-(8S-13) |K| of a whole factorisation (S = 6: 2.9e-11; measured < 2e-11 at 1,100-1,500 leaves).
double drp_ozaki_cond_limit()
{
    static double v = -1.0;
    if (v < 0.0) {
        const char* e = getenv("DRP_OZAKI_COND_LIMIT");
        const int S = drp_ozaki_slices();
        v = e ? atof(e) : 1e-5 * ldexp(1.0, 8 * (S ? S : 6) - 13);
    }
    return v;
}

int drp_ozaki_slices()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_OZAKI_SLICES");
        v = e ? atoi(e) : 6;
        if (v != 0 && (v < 4 || v > 7)) v = 6;
    }
    return v;
}

// Where the slices of the panel P = M[c1 .. R) x [k0, k0 + K) live inside the workspace (the producer may be oz_slice_kernel
// or, for the 64-column steps inside a panel, the trsm kernel that has just solved those columns: factor.cu).  Returns false
// when an update of this shape does not take the tensor path at all.
bool drp_ozaki_layout(int z, int K, int c1, int R, int col_hi, int S, void* ws, int64_t ws_bytes, int buffer, DrpOzLayout* out)
{
    if (S == 0 || !ws) return false;
    if (K < 32 || K > 256) return false;
    const int KBN = K > 128 ? 2 : 1;
    if (S * KBN > 12) return false;
    const int cmax = R < col_hi ? R : col_hi;
    if (R - c1 < 384 || cmax - c1 < oz_min_cols()) return false;
    const int rb = c1 & ~7;
    const int nTi = (R - rb + OZ_TM - 1) / OZ_TM;
    out->rb = rb;
    out->Rpad = nTi * OZ_TM;
    out->KBN = KBN;
    out->sl_bs = (int64_t)S * KBN * out->Rpad * OZ_ROWB;
    const int64_t need = (int64_t)z * out->sl_bs + (int64_t)z * out->Rpad * 8 + 2048;
    const int64_t bbytes = (ws_bytes / 2) & ~(int64_t)1023;      // two buffers
    if (need > bbytes - 1024) return false;
    uint8_t* w = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023) + (buffer ? bbytes : 0);
    out->sl = w;
    out->scl = reinterpret_cast<int32_t*>(w + (int64_t)z * out->sl_bs);
    out->counter = reinterpret_cast<int*>(w + (int64_t)z * out->sl_bs + (int64_t)z * out->Rpad * 8);
    return true;
}

// Trailing update through the int8 tensor pipe.  Returns 0 when it ran, 1 when the shape/workspace does not qualify
// (the caller then uses the fp64 mma.sync kernel), otherwise a launch error code.
//   the panel P = M[c1 .. R) x [k0, k0 + K) is sliced from row c1 down (do_slice; `buffer` 0/1 selects the slice buffer);
//   updated entries: rows r in [c1, R), columns c in [col_lo, min(R, col_hi)), c <= r  (col_lo >= c1).
//   do_slice = 0 re-uses the slices a previous call with the same (c1, R, k0, K, buffer) made: the second column-range part
//   of one update, possibly running on another stream (counter_idx 1).
int drp_ozaki_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_lo, int col_hi, int S,
                   void* ws, int64_t ws_bytes, const int* zmap, int do_slice, int buffer, int counter_idx, cudaStream_t st)
{
    DrpOzLayout lay;
    if (!drp_ozaki_layout(z, K, c1, R, col_hi, S, ws, ws_bytes, buffer, &lay)) return 1;
    const int KBN = lay.KBN;
    const int cmax = R < col_hi ? R : col_hi;
    if (col_lo < c1) col_lo = c1;
    const int below = R - c1, cl = cmax - col_lo;
    if (cl <= 0) return 0;
    OzParams p;
    p.M = M;
    p.ld = ld;
    p.bs = bs;
    p.rb = lay.rb;
    p.c1 = col_lo;                                       // rows and columns below col_lo are not touched (c <= r)
    p.R = R;
    p.cmax = cmax;
    p.K = K;
    p.trace = g_oz_trace;
    p.prof = g_oz_prof;
    p.zmap = zmap;
    p.nTi = lay.Rpad / OZ_TM;
    p.nTj = (cmax - p.rb + OZ_TN - 1) / OZ_TN;
    p.ti0 = (col_lo - p.rb) / OZ_TM;
    p.tj0 = (col_lo - p.rb) / OZ_TN;
    p.Rpad = lay.Rpad;
    p.sl_bs = lay.sl_bs;
    p.sl = lay.sl;
    p.scl = lay.scl;
    p.counter_idx = counter_idx ? 1 : 0;
    p.counter = lay.counter + p.counter_idx;
    // NOTE: the slicing kernel writes zeros for the rows of the panel outside [c1, R); it is given the panel's own c1
    int upz = 0;
    for (int ti = p.ti0; ti < p.nTi; ++ti) {
        const int tot = (2 * ti + 2 < p.nTj ? 2 * ti + 2 : p.nTj) - p.tj0;
        if (tot > 0) upz += (tot + OZ_CHUNK - 1) / OZ_CHUNK;
    }
    if (upz == 0) return 0;
    p.units_per_z = upz;
    p.total_units = upz * z;
    const double c0 = (double)(col_lo - c1), c1d = (double)(cmax - c1);      // columns [c0, c1d) of the trapezoid below c1
    const double entries = (double)below * (c1d - c0) - (c1d * (c1d - 1.0) - c0 * (c0 - 1.0)) / 2.0;
    const double flops = (double)z * entries * 2.0 * K;
    p.slice_c1 = c1;
    int rc = 1;
    if (KBN == 1) {
        switch (S) {
            case 4: rc = oz_launch<4, 1>(p, z, k0, K, st, flops, do_slice != 0); break;
            case 5: rc = oz_launch<5, 1>(p, z, k0, K, st, flops, do_slice != 0); break;
            case 6: rc = oz_launch<6, 1>(p, z, k0, K, st, flops, do_slice != 0); break;
            case 7: rc = oz_launch<7, 1>(p, z, k0, K, st, flops, do_slice != 0); break;
        }
    } else {
        switch (S) {
            case 4: rc = oz_launch<4, 2>(p, z, k0, K, st, flops, do_slice != 0); break;
            case 5: rc = oz_launch<5, 2>(p, z, k0, K, st, flops, do_slice != 0); break;
            case 6: rc = oz_launch<6, 2>(p, z, k0, K, st, flops, do_slice != 0); break;
        }
    }
    if (rc) return rc;
    return drp_launch_status();
}
