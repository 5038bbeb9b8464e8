// Bidirectional single-layer GRU recurrence (hidden size 60), forward and backward through time, fp64.
//
// Replaces nn.GRU inside the reference's decoder and encoder (S/ = /root/reference/draupnir/src/draupnir/):
//   RNNDecoder_Tiling.forward      S/models_utils.py:93-99    rnn_output, _ = self.rnn(input, hidden)
//   RNNEncoder.forward             S/models_utils.py:52-65    rnn_hidden_states, rnn_final = self.rnn(input, hidden)
// with torch's gate conventions (gate order r, z, n):
//   r = sigmoid(gi_r + W_hr h + b_hr)      z = sigmoid(gi_z + W_hz h + b_hz)
//   n = tanh(gi_n + r * (W_hn h + b_hn))   h' = (1 - z) n + z h
// `gi = W_ih x + b_ih` is the caller's business (one dense GEMM for the encoder; for the decoder the input is
// [latent_i ; blosum_l], so gi[i,l] = U[i] + V[l] is a broadcast sum and the [n,L,51] input is never materialised).
//
// The recurrence is the sequential part: L dependent steps of a [rows,60]x[60,180] product plus gate math.  One CTA
// owns 32 sequences of one direction for all L steps:
//   * W_hh lives in REGISTERS as mma.sync.m8n8k4.f64 B-fragments (45 doubles per lane), h in shared memory;
//   * warp w owns hidden units 8w..8w+7 of all three gates and all 32 rows, so every thread ends up holding r, z, n
//     pre-activations of the SAME (row, unit) pairs and the gate math needs no exchange;
//   * the per-step operand rows (gi; in the backward pass the saved gates and h_prev) are contiguous per sequence and
//     are streamed into a shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier), issued one step ahead.
// The backward kernel back-propagates dh through time (dh_prev = dh z + dgh W_hh, again register-resident W_hh) and
// writes dgi; the weight gradients are dense reductions over all (sequence, step) pairs and are left to one GEMM each.
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int GH = 60;              // hidden size
constexpr int G3 = 3 * GH;          // gate columns
constexpr int GHS = 68;             // smem row stride of h   (stride mod 16 == 4: conflict-free fragment loads)
constexpr int GDS = 180;            // smem row stride of dgh (180 mod 16 == 4)
constexpr int GTHREADS = 256;
constexpr long long G_TIMEOUT_CYCLES = 20000000000LL;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t g_mbar_try_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity)
{
    if (g_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    int spins = 0;
    while (!g_mbar_try_wait(bar, parity)) {
        if (((++spins) & 0xFFF) == 0 && clock64() - t0 > G_TIMEOUT_CYCLES) __trap();
    }
}
__device__ __forceinline__ void g_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// 1 / x for x in [1, 1e304] (the denominators 1 + e^y below): MUFU.RCP64H seed (~2^-19) + two Newton steps, <= 1 ulp.
// Branch-free on purpose: __drcp_rn carries a slow-path CALL that keeps the compiler from interleaving the gate
// arithmetic of a thread's 2 MT independent elements, which made the gate phase a serial latency chain
// (profiles/r01/trace_gru_v1.txt: 8.8 k of the 15.5 k cycles of a 32-row step).
__device__ __forceinline__ double gru_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double sigmoid_f64(double x) { return gru_rcp(1.0 + drp_exp(-x)); }
// tanh(x) = 1 - 2 / (1 + e^{2x}): absolute error ~1e-16 (what the recurrence needs), saturates correctly at +-1
__device__ __forceinline__ double tanh_f64(double x) { return fma(-2.0, gru_rcp(1.0 + drp_exp(2.0 * x)), 1.0); }

// Lean variants for the second-generation forward kernel: about half the instructions of the pair above (the gate phase is
// bound by instruction issue and dependent-issue latency, not by the fp64 pipe).  Both work on a = min(|x|, 704) — two
// integer instructions on the high word instead of a two-sided fp64 clamp — evaluate e^{-a} in (0, 1] (no overflow case,
// denominators in (1, 2]) with the exponent table in shared memory, and restore the sign at the end.
__device__ __forceinline__ double gru_abs_clamped(double x)
{
    return __hiloint2double(min(__double2hiint(x) & 0x7fffffff, 0x40860000), __double2loint(x));
}
__device__ __forceinline__ double gru_copysign_pos(double m, double x)      // m >= 0
{
    return __hiloint2double(__double2hiint(m) | (__double2hiint(x) & 0x80000000), __double2loint(m));
}
__device__ __forceinline__ double gru_expm(double a, const double* __restrict__ tab)     // e^{-a}, 0 <= a <= 1408
{
    const double sh = fma(a, -46.166241308446828384, 6755399441055744.0);   // rint(-a 32/ln2) by the 1.5 * 2^52 shift
    const int n = max(__double2loint(sh), -32000);                          // 2^-1000: far below 1 ulp of the results
    const double t = sh - 6755399441055744.0;
    double r = fma(t, -0.021660849219188094, -a);
    r = fma(t, -1.733101967801894e-10, r);
    double q = 1.3888888888888889e-03;
    q = fma(q, r, 8.3333333333333332e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double T = tab[n & 31];
    const double v = fma(T, q * r, T);
    return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
}
__device__ __forceinline__ double gru_sigmoid(double x, const double* __restrict__ tab)
{
    const double y = gru_rcp(1.0 + gru_expm(gru_abs_clamped(x), tab));      // sigma(|x|) in [1/2, 1)
    return 0.5 + gru_copysign_pos(y - 0.5, x);
}
__device__ __forceinline__ double gru_tanh(double x, const double* __restrict__ tab)
{
    const double a = gru_abs_clamped(x);
    const double y = gru_rcp(1.0 + gru_expm(a + a, tab));
    return gru_copysign_pos(fma(2.0, y, -1.0), x);
}

// ---------------------------------------------------------------------------------------------------------------
// Forward.  grid (ceil(n/ROWS), 2 directions), 256 threads; ROWS = 8 MT sequences per CTA (MT = 4, 2 or 1 chosen by
// the host so that one wave of CTAs covers the batch: the recurrence is latency-bound, fewer rows = shorter steps).
//   gi    [n, L, 2, 180]   W_ih x + b_ih for both directions
//   whh   [2, 180, 60]     bhh [2, 180]     h0 [2, n, 60]
//   out   [n, L, 120]      hidden states (forward direction in [:60], backward in [60:], torch layout)
//   gates [n, L, 2, 4, 60] r, z, n, (W_hn h + b_hn) for the backward pass; nullable
// ---------------------------------------------------------------------------------------------------------------
#ifdef DRP_GRU_TRACE
// profiling build only (profiles/trace_gru.py): clock64 at the phase boundaries of the forward step, CTA (0,0), warps 0 and 8 (7 in 256-thread CTAs)
__device__ long long g_gru_trace[8 * 64 * 2];
#define GRU_TRACE(slot)                                                                                   \
    do {                                                                                                  \
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == (blockDim.x > 256 ? 8 : (blockDim.x == 160 ? 3 : 7))) && t >= 200 && t < 264) \
            g_gru_trace[((warp ? 1 : 0) * 64 + (t - 200)) * 8 + (slot)] = clock64();                     \
    } while (0)
extern "C" __attribute__((visibility("default"))) int drp_gru_trace_read(long long* host)
{
    return (int)cudaMemcpyFromSymbol(host, g_gru_trace, sizeof(g_gru_trace));
}
#define GRU_TRACE_B(slot)                                                                                 \
    do {                                                                                                  \
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == (blockDim.x == 160 ? 3 : 7)) && tb >= 200 && tb < 264) \
            g_gru_trace[((warp ? 1 : 0) * 64 + (tb - 200)) * 8 + (slot)] = clock64();                    \
    } while (0)
#define GRU_TRACE_W(slot)                                                                                 \
    do {                                                                                                  \
        if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7) && it >= 200 && it < 264)            \
            g_gru_trace[((warp ? 1 : 0) * 64 + (it - 200)) * 8 + (slot)] = clock64();                    \
    } while (0)
#else
#define GRU_TRACE(slot) do { } while (0)
#define GRU_TRACE_B(slot) do { } while (0)
#define GRU_TRACE_W(slot) do { } while (0)
#endif
constexpr int GF_STAGES = 3;
__host__ __device__ constexpr int gf_smem(int rows, bool uv)        // h double buffer + gi ring + barriers (+ the CTA's U rows)
{
    return 2 * rows * GHS * 8 + GF_STAGES * rows * G3 * 8 + 64 + (uv ? rows * G3 * 8 : 0);
}

// UV = true: gi[i][l] = U[i] + V[l] (the decoder's [latent_i ; blosum_l] input, S/models.py:339-342): `gi` then points to
// U [n, 2, 180] and `vrow` to V [L, 2, 180]; the U rows of the CTA stay in shared memory, one V row streams per step.
template <bool UV, int MT>
__global__ void __launch_bounds__(GTHREADS, 1) gru_fwd_kernel(const double* __restrict__ gi, const double* __restrict__ vrow,
                                                              const double* __restrict__ whh,
                                                              const double* __restrict__ bhh, const double* __restrict__ h0,
                                                              int n, int L, double* __restrict__ out,
                                                              double* __restrict__ gates, int gi_dirs)
{
    constexpr int GROWS = 8 * MT;
    constexpr int GF_STAGE_BYTES = GROWS * G3 * 8;
    extern __shared__ __align__(16) uint8_t gsm[];
    double* hbuf = reinterpret_cast<double*>(gsm);                                   // [2][32][68]
    double* ring = reinterpret_cast<double*>(gsm + 2 * GROWS * GHS * 8);              // [stages][32][180]
    const uint32_t bars = g_smem_u32(gsm + 2 * GROWS * GHS * 8 + GF_STAGES * GF_STAGE_BYTES);
    double* ubuf = reinterpret_cast<double*>(gsm + gf_smem(GROWS, false));             // [rows][180] (UV only)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * GROWS;
    const int nrows = min(GROWS, n - row0);
    const int j0 = warp * 8 + 2 * tig;            // this thread's hidden units j0, j0+1 (as C-fragment columns)
    const bool jok = j0 < GH;
    const int jb = warp * 8 + gid;                // hidden unit whose W_hh row this lane holds as B-fragment
    const double* W = whh + (int64_t)dir * G3 * GH;

    double B[3][15];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int ks = 0; ks < 15; ++ks) B[g][ks] = jb < GH ? W[(int64_t)(g * GH + jb) * GH + 4 * ks + tig] : 0.0;
    double bh[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bh[g][0] = jok ? bhh[dir * G3 + g * GH + j0] : 0.0;
        bh[g][1] = jok ? bhh[dir * G3 + g * GH + j0 + 1] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < GF_STAGES; ++s) g_mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // h0 -> hbuf[0] (all 60 units of the CTA's rows) and the thread's own copies
    for (int idx = tid; idx < GROWS * GH; idx += GTHREADS) {
        const int r = idx / GH, j = idx - r * GH;
        hbuf[r * GHS + j] = r < nrows ? h0[((int64_t)dir * n + row0 + r) * GH + j] : 0.0;
    }
    if (UV) {
        for (int idx = tid; idx < GROWS * G3; idx += GTHREADS) {
            const int r = idx / G3, c = idx - r * G3;
            ubuf[idx] = r < nrows ? gi[((int64_t)(row0 + r) * 2 + dir) * G3 + c] : 0.0;
        }
    }
    __syncthreads();
    double hold[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        hold[mt][0] = jok ? hbuf[(8 * mt + gid) * GHS + j0] : 0.0;
        hold[mt][1] = jok ? hbuf[(8 * mt + gid) * GHS + j0 + 1] : 0.0;
    }
    auto issue = [&](int t) {      // warp 0: stream gi of step t into its ring stage (one bulk copy per sequence)
        const int s = t % GF_STAGES;
        const int l = dir ? L - 1 - t : t;
        if (UV) {
            if (lane == 0) {
                g_mbar_expect_tx(bars + 8 * s, G3 * 8);
                g_bulk_g2s(g_smem_u32(ring + (int64_t)s * GROWS * G3), vrow + ((int64_t)l * 2 + dir) * G3, G3 * 8, bars + 8 * s);
            }
            return;
        }
        // every warp issues the copies of MT rows (a bulk-copy issue costs ~60 cycles; 32 of them on one warp put that warp
        // 1.9 k cycles behind the others at every step); the single arrival that carries the byte count is warp 0's, so
        // the phase cannot complete before it, whatever the order in which the copies land
        if (warp == 0 && lane == 0) g_mbar_expect_tx(bars + 8 * s, (uint32_t)nrows * G3 * 8);
        const int rr = warp * MT + lane;
        if (lane < MT && rr < nrows)
            g_bulk_g2s(g_smem_u32(ring + (int64_t)s * GROWS * G3 + rr * G3),
                       gi + (((int64_t)(row0 + rr) * L + l) * gi_dirs + (gi_dirs == 2 ? dir : 0)) * G3, G3 * 8, bars + 8 * s);
    };
    if (!UV || warp == 0)
        for (int t = 0; t < min(GF_STAGES - 1, L); ++t) issue(t);

    for (int t = 0; t < L; ++t) {
        const int l = dir ? L - 1 - t : t;
        const double* hc = hbuf + (t & 1) * GROWS * GHS;
        double* hn = hbuf + ((t + 1) & 1) * GROWS * GHS;
        GRU_TRACE(0);
        if ((!UV || warp == 0) && t + GF_STAGES - 1 < L) issue(t + GF_STAGES - 1);   // its stage was drained in step t-1
        GRU_TRACE(1);
        double acc[3][MT][2];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) acc[g][mt][0] = acc[g][mt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 15; ++ks) {
            double a[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a[mt] = hc[(8 * mt + gid) * GHS + 4 * ks + tig];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma_8x8x4(acc[g][mt][0], acc[g][mt][1], a[mt], B[g][ks]);
        }
        const int s = t % GF_STAGES;
        GRU_TRACE(2);
        g_mbar_wait(bars + 8 * s, (uint32_t)(t / GF_STAGES) & 1u);
        GRU_TRACE(3);
        const double* gs = ring + (int64_t)s * GROWS * G3;
        if (jok) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r = 8 * mt + gid;
                double2 xr, xz, xn;
                if (UV) {
                    const double2 ur = *reinterpret_cast<const double2*>(ubuf + r * G3 + j0);
                    const double2 uz = *reinterpret_cast<const double2*>(ubuf + r * G3 + GH + j0);
                    const double2 un = *reinterpret_cast<const double2*>(ubuf + r * G3 + 2 * GH + j0);
                    const double2 vr = *reinterpret_cast<const double2*>(gs + j0);
                    const double2 vz = *reinterpret_cast<const double2*>(gs + GH + j0);
                    const double2 vn = *reinterpret_cast<const double2*>(gs + 2 * GH + j0);
                    xr = make_double2(ur.x + vr.x, ur.y + vr.y);
                    xz = make_double2(uz.x + vz.x, uz.y + vz.y);
                    xn = make_double2(un.x + vn.x, un.y + vn.y);
                } else {
                    xr = *reinterpret_cast<const double2*>(gs + r * G3 + j0);
                    xz = *reinterpret_cast<const double2*>(gs + r * G3 + GH + j0);
                    xn = *reinterpret_cast<const double2*>(gs + r * G3 + 2 * GH + j0);
                }
                double hv[2], rv[2], zv[2], nv[2], qv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double gr = (e ? xr.y : xr.x) + acc[0][mt][e] + bh[0][e];
                    const double gz = (e ? xz.y : xz.x) + acc[1][mt][e] + bh[1][e];
                    qv[e] = acc[2][mt][e] + bh[2][e];
                    rv[e] = sigmoid_f64(gr);
                    zv[e] = sigmoid_f64(gz);
                    nv[e] = tanh_f64((e ? xn.y : xn.x) + rv[e] * qv[e]);
                    hv[e] = (1.0 - zv[e]) * nv[e] + zv[e] * hold[mt][e];
                    hold[mt][e] = hv[e];
                }
                *reinterpret_cast<double2*>(hn + r * GHS + j0) = make_double2(hv[0], hv[1]);
                if (r < nrows) {
                    const int64_t site = (int64_t)(row0 + r) * L + l;
                    *reinterpret_cast<double2*>(out + site * (2 * GH) + dir * GH + j0) = make_double2(hv[0], hv[1]);
                    if (gates) {
                        double* gp = gates + (site * 2 + dir) * (4 * GH) + j0;
                        *reinterpret_cast<double2*>(gp) = make_double2(rv[0], rv[1]);
                        *reinterpret_cast<double2*>(gp + GH) = make_double2(zv[0], zv[1]);
                        *reinterpret_cast<double2*>(gp + 2 * GH) = make_double2(nv[0], nv[1]);
                        *reinterpret_cast<double2*>(gp + 3 * GH) = make_double2(qv[0], qv[1]);
                    }
                }
            }
        }
        GRU_TRACE(4);
        __syncthreads();      // h of step t complete in hn; ring stage s and hc free for reuse
        GRU_TRACE(5);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward, second generation: W_hh lives in SHARED memory (k-major, row stride 184: conflict-free B-fragment loads) instead
// of 90 registers per thread, which (a) lets 16 warps share a CTA (warp = hidden-unit group x row half) so that the
// latency-bound gate arithmetic has twice the warps to hide behind, and (b) leaves the compiler the registers to
// interleave the independent elements of a thread (profiles/r01/trace_gru_v2.txt: at 32 rows the gate phase took 6.5-7.1 k
// of the 13.7 k cycles of a step although its fp64 work is ~2 k cycles).  Same arguments and layouts as gru_fwd_kernel.
//   WS = 2: 512 threads, each warp owns MT/2 row tiles;  WS = 1: 256 threads (8-row CTAs).
// ---------------------------------------------------------------------------------------------------------------
constexpr int GWS = 180;      // row stride of W^T in shared memory: 180 mod 16 == 4, so the 16 lanes of a half-warp (gid 0-3 x tig 0-3,
                              // address tig * GWS + gid) hit 16 distinct 8-byte banks
constexpr int GRS = 184;      // row stride of the gi / U rows: 16-byte accesses at r * GRS + 2 tig want GRS / 2 == 4 (mod 8)
constexpr int GF2_STAGES = 2;
__host__ __device__ constexpr int gf2_smem(int rows, bool uv)
{
    // h double buffer | W^T | per-half rings (UV: 2 halves x stages x one V row, + the CTA's U rows) | barriers
    return 2 * rows * GHS * 8 + (GH * GWS + 8) * 8 + (uv ? (2 * GF2_STAGES * GRS * 8 + rows * GRS * 8) : GF2_STAGES * rows * GRS * 8) +
           64 + 256;                                   // + the 32-entry 2^{j/32} table of the exponential
}
__device__ __forceinline__ void g_named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void g_named_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// The two row halves of a 512-thread CTA (warps 0-7 / 8-15) never exchange data: each has its own ring, mbarriers and named
// barrier.  They take turns on the recurrent product (token passed through named barriers 3/4), so that one half's gate
// arithmetic (latency/issue-bound, fp64 pipe ~1/3 busy) runs under the other half's product (fp64 pipe saturated by DMMA).
// Left alone the halves fall back into lockstep (the trailing half gets the whole pipe and catches up), hence the token.
template <bool UV, int MT, int WS>
__global__ void __launch_bounds__(256 * WS, 1) gru_fwd2_kernel(const double* __restrict__ gi, const double* __restrict__ vrow,
                                                               const double* __restrict__ whh,
                                                               const double* __restrict__ bhh, const double* __restrict__ h0,
                                                               int n, int L, double* __restrict__ out,
                                                               double* __restrict__ gates, int gi_dirs, int stagger_ns)
{
    constexpr int NT = 256 * WS;
    constexpr int GROWS = 8 * MT;
    constexpr int MTW = MT / WS;                                      // row tiles per warp
    constexpr int HROWS = 8 * MTW;                                    // rows per half
    constexpr int STAGE_ROWS = UV ? 1 : HROWS;
    static_assert(MT % WS == 0, "row tiles must split evenly over the warp halves");
    extern __shared__ __align__(16) uint8_t gsm[];
    double* hbuf = reinterpret_cast<double*>(gsm);                                   // [2][rows][68]
    double* wt = hbuf + 2 * GROWS * GHS;                                              // [60][184]: wt[k][c] = W_hh[c][k]
    double* ring0 = wt + GH * GWS + 8;                                                // [half][stage][rows/half or 1][184]
    double* ubuf = ring0 + WS * GF2_STAGES * STAGE_ROWS * GRS + (UV && WS == 1 ? GF2_STAGES * GRS : 0);   // [rows][184] (UV only)
    double* etab = reinterpret_cast<double*>(gsm + gf2_smem(GROWS, UV) - 256);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int wg = warp & 7, wr = warp >> 3;                                          // hidden-unit group, row half
    double* ring = ring0 + wr * GF2_STAGES * STAGE_ROWS * GRS;
    const uint32_t bars = g_smem_u32(gsm + gf2_smem(GROWS, UV) - 320) + wr * GF2_STAGES * 8;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * GROWS;
    const int nrows = min(GROWS, n - row0);
    const int hrow0 = wr * HROWS;                                                     // first CTA row of this half
    const int nrows_h = max(0, min(HROWS, nrows - hrow0));
    const int j0 = wg * 8 + 2 * tig;              // this thread's hidden units j0, j0+1 (as C-fragment columns)
    const bool jok = j0 < GH;
    const int jb = wg * 8 + gid;                  // hidden unit whose W_hh row this lane feeds as B-fragment
    const double* W = whh + (int64_t)dir * G3 * GH;

    for (int idx = tid; idx < GH * GWS + 8; idx += NT) {     // (+8: the unused lanes of hidden group 7 read up to 3 past the end)
        const int k = idx / GWS, c = idx - k * GWS;
        wt[idx] = k < GH ? W[(int64_t)c * GH + k] : 0.0;
    }
    if (tid < 32) etab[tid] = DRP_EXP2_TAB[tid];
    double bh[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bh[g][0] = jok ? bhh[dir * G3 + g * GH + j0] : 0.0;
        bh[g][1] = jok ? bhh[dir * G3 + g * GH + j0 + 1] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < WS * GF2_STAGES; ++s) g_mbar_init(g_smem_u32(gsm + gf2_smem(GROWS, UV) - 320) + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int idx = tid; idx < GROWS * GH; idx += NT) {
        const int r = idx / GH, j = idx - r * GH;
        hbuf[r * GHS + j] = r < nrows ? h0[((int64_t)dir * n + row0 + r) * GH + j] : 0.0;
    }
    if (UV) {
        for (int idx = tid; idx < GROWS * G3; idx += NT) {
            const int r = idx / G3, c = idx - r * G3;
            ubuf[r * GRS + c] = r < nrows ? gi[((int64_t)(row0 + r) * 2 + dir) * G3 + c] : 0.0;
        }
    }
    __syncthreads();
    double hold[MTW][2];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
        const int r = hrow0 + 8 * mt + gid;
        hold[mt][0] = jok ? hbuf[r * GHS + j0] : 0.0;
        hold[mt][1] = jok ? hbuf[r * GHS + j0 + 1] : 0.0;
    }
    auto issue = [&](int t) {      // stream the gi rows (UV: the V row) of step t into this half's ring stage
        const int s = t % GF2_STAGES;
        const int l = dir ? L - 1 - t : t;
        if (UV) {
            if (wg == 0 && lane == 0) {
                g_mbar_expect_tx(bars + 8 * s, G3 * 8);
                g_bulk_g2s(g_smem_u32(ring + s * GRS), vrow + ((int64_t)l * 2 + dir) * G3, G3 * 8, bars + 8 * s);
            }
            return;
        }
        // the half's first warp carries the byte count, so the phase cannot complete before it whatever the landing order
        if (wg == 0 && lane == 0) g_mbar_expect_tx(bars + 8 * s, (uint32_t)nrows_h * G3 * 8);
        const int rr = wg * MTW + lane;                                 // MTW copies per warp (an issue costs ~60 cycles)
        if (lane < MTW && rr < nrows_h)
            g_bulk_g2s(g_smem_u32(ring + (int64_t)s * HROWS * GRS + rr * GRS),
                       gi + (((int64_t)(row0 + hrow0 + rr) * L + l) * gi_dirs + (gi_dirs == 2 ? dir : 0)) * G3, G3 * 8,
                       bars + 8 * s);
    };
    issue(0);
    const bool handoff = WS == 2 && stagger_ns != 0;

    for (int t = 0; t < L; ++t) {
        const int l = dir ? L - 1 - t : t;
        const double* hc = hbuf + (t & 1) * GROWS * GHS;
        double* hn = hbuf + ((t + 1) & 1) * GROWS * GHS;
        GRU_TRACE(0);
        if (t + 1 < L) issue(t + 1);               // the other stage was drained before the barrier that ended step t-1
        if (handoff && (wr == 1 || t > 0)) g_named_barrier(3 + wr, 512);   // the other half has finished its product
        GRU_TRACE(1);
        double acc[3][MTW][2];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) acc[g][mt][0] = bh[g][0], acc[g][mt][1] = bh[g][1];
#pragma unroll
        for (int ks = 0; ks < 15; ++ks) {
            double a[MTW], b[3];
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) a[mt] = hc[(hrow0 + 8 * mt + gid) * GHS + 4 * ks + tig];
#pragma unroll
            for (int g = 0; g < 3; ++g) b[g] = wt[(4 * ks + tig) * GWS + g * GH + jb];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) dmma_8x8x4(acc[g][mt][0], acc[g][mt][1], a[mt], b[g]);
        }
        if (handoff && (wr == 0 || t + 1 < L)) g_named_arrive(3 + (wr ^ 1), 512);   // the product token goes to the other half
        const int s = t % GF2_STAGES;
        GRU_TRACE(2);
        g_mbar_wait(bars + 8 * s, (uint32_t)(t / GF2_STAGES) & 1u);
        GRU_TRACE(3);
        const double* gs = ring + (int64_t)s * STAGE_ROWS * GRS;
        if (jok) {
            double xr[MTW][2], xz[MTW][2], xn[MTW][2];
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                const int r = 8 * mt + gid;                              // row within the half
                double2 ar, az, an;
                if (UV) {
                    const double* up = ubuf + (hrow0 + r) * GRS;
                    const double2 ur = *reinterpret_cast<const double2*>(up + j0);
                    const double2 uz = *reinterpret_cast<const double2*>(up + GH + j0);
                    const double2 un = *reinterpret_cast<const double2*>(up + 2 * GH + j0);
                    const double2 vr = *reinterpret_cast<const double2*>(gs + j0);
                    const double2 vz = *reinterpret_cast<const double2*>(gs + GH + j0);
                    const double2 vn = *reinterpret_cast<const double2*>(gs + 2 * GH + j0);
                    ar = make_double2(ur.x + vr.x, ur.y + vr.y);
                    az = make_double2(uz.x + vz.x, uz.y + vz.y);
                    an = make_double2(un.x + vn.x, un.y + vn.y);
                } else {
                    ar = *reinterpret_cast<const double2*>(gs + r * GRS + j0);
                    az = *reinterpret_cast<const double2*>(gs + r * GRS + GH + j0);
                    an = *reinterpret_cast<const double2*>(gs + r * GRS + 2 * GH + j0);
                }
                xr[mt][0] = ar.x, xr[mt][1] = ar.y, xz[mt][0] = az.x, xz[mt][1] = az.y, xn[mt][0] = an.x, xn[mt][1] = an.y;
            }
            double hv[MTW][2], rv[MTW][2], zv[MTW][2], nv[MTW][2], qv[MTW][2];
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    qv[mt][e] = acc[2][mt][e];                           // the accumulators started from b_hh
                    xr[mt][e] += acc[0][mt][e];
                    xz[mt][e] += acc[1][mt][e];
                    rv[mt][e] = gru_sigmoid(xr[mt][e], etab);
                    zv[mt][e] = gru_sigmoid(xz[mt][e], etab);
                }
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double gn = fma(rv[mt][e], qv[mt][e], xn[mt][e]);
                    nv[mt][e] = gru_tanh(gn, etab);
                    hv[mt][e] = fma(zv[mt][e], hold[mt][e] - nv[mt][e], nv[mt][e]);
                    // the integer clamp of the lean gate functions hides a NaN pre-activation: hand it on to h (0 * finite = 0)
                    hv[mt][e] = fma(0.0, (xr[mt][e] + xz[mt][e]) + gn, hv[mt][e]);
                    hold[mt][e] = hv[mt][e];
                }
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                const int r = hrow0 + 8 * mt + gid;
                *reinterpret_cast<double2*>(hn + r * GHS + j0) = make_double2(hv[mt][0], hv[mt][1]);
                if (r < nrows) {
                    const int64_t site = (int64_t)(row0 + r) * L + l;
                    *reinterpret_cast<double2*>(out + site * (2 * GH) + dir * GH + j0) = make_double2(hv[mt][0], hv[mt][1]);
                    if (gates) {
                        double* gp = gates + (site * 2 + dir) * (4 * GH) + j0;
                        *reinterpret_cast<double2*>(gp) = make_double2(rv[mt][0], rv[mt][1]);
                        *reinterpret_cast<double2*>(gp + GH) = make_double2(zv[mt][0], zv[mt][1]);
                        *reinterpret_cast<double2*>(gp + 2 * GH) = make_double2(nv[mt][0], nv[mt][1]);
                        *reinterpret_cast<double2*>(gp + 3 * GH) = make_double2(qv[mt][0], qv[mt][1]);
                    }
                }
            }
        }
        GRU_TRACE(4);
        g_named_barrier(1 + wr, 256);      // this half: h of step t complete in hn; its ring stage s and hc rows free for reuse
        GRU_TRACE(5);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward through time.  Steps run in the reverse of the forward order of the direction.
//   dout  [n, L, 120]  gradient w.r.t. the hidden states   out / gates / h0 / whh as in the forward kernel
//   dgi   [n, L, 2, 180]  gradient w.r.t. gi (r, z, n pre-activations); dgh = dgi with the n-part multiplied by r
//   dh0   [2, n, 60] nullable
// ---------------------------------------------------------------------------------------------------------------
constexpr int GB_STAGES = 2;
constexpr int GB_ROW_BYTES = (4 * GH + GH) * 8;                                   // saved gates + h_prev = 2400
constexpr int GB_ROW_STRIDE = 312;     // doubles between ring rows: 16-byte accesses at r * stride + 2 tig want stride / 2 == 4 (mod 8)
__host__ __device__ constexpr int gb_smem(int rows) { return rows * GDS * 8 + GB_STAGES * rows * GB_ROW_STRIDE * 8 + 64; }

template <int MT>
__global__ void __launch_bounds__(GTHREADS, 1) gru_bwd_kernel(const double* __restrict__ dout, const double* __restrict__ out,
                                                              const double* __restrict__ gates, const double* __restrict__ whh,
                                                              const double* __restrict__ h0, int n, int L,
                                                              double* __restrict__ dgi, double* __restrict__ dh0,
                                                              int dgi_dirs, int dout_last)
{
    // dout_last != 0: `dout` is [n, 120] and holds the gradient of the hidden state at the LAST position only (every
    // other position's gradient is zero): the encoder's case, which spares a zero-filled [n, L, 120] tensor
    constexpr int GROWS = 8 * MT;
    constexpr int GB_STAGE_BYTES = GROWS * GB_ROW_STRIDE * 8;
    extern __shared__ __align__(16) uint8_t gsm[];
    double* dgh = reinterpret_cast<double*>(gsm);                                    // [32][180]
    double* ring = reinterpret_cast<double*>(gsm + GROWS * GDS * 8);                  // [stages][32][300]
    const uint32_t bars = g_smem_u32(gsm + GROWS * GDS * 8 + GB_STAGES * GB_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * GROWS;
    const int nrows = min(GROWS, n - row0);
    const int j0 = warp * 8 + 2 * tig;
    const bool jok = j0 < GH;
    const int jb = warp * 8 + gid;
    const double* W = whh + (int64_t)dir * G3 * GH;
    constexpr int RS = GB_ROW_STRIDE;             // ring row: [r | z | n | q | h_prev] = 300 doubles, padded to 312

    double B[45];                                 // B[k][n] = W_hh[c = 4 ks + tig][j = jb]
#pragma unroll
    for (int ks = 0; ks < 45; ++ks) B[ks] = jb < GH ? W[(int64_t)(4 * ks + tig) * GH + jb] : 0.0;
    if (tid == 0) {
        for (int s = 0; s < GB_STAGES; ++s) g_mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int idx = tid; idx < GROWS * GDS; idx += GTHREADS) dgh[idx] = 0.0;
    __syncthreads();
    // backward step tb = 0..L-1 handles forward step t = L-1-tb, i.e. position l = dir ? tb : L-1-tb
    auto issue = [&](int tb) {
        const int s = tb % GB_STAGES;
        const int t = L - 1 - tb;
        const int l = dir ? L - 1 - t : t;
        const int lp = dir ? l + 1 : l - 1;       // position of h_prev (outside [0,L): h0)
        // all warps issue (2 MT copies each, one per lane): 64 bulk-copy issues on one warp cost it ~3.8 k cycles per step.
        // Warp 0's arrival carries the byte count, so the phase cannot complete early whatever the landing order.
        if (warp == 0 && lane == 0) g_mbar_expect_tx(bars + 8 * s, (uint32_t)nrows * GB_ROW_BYTES);
        const int rr = warp * MT + (lane >> 1);
        if (lane < 2 * MT && rr < nrows) {
            const int64_t row = row0 + rr;
            double* dst = ring + (int64_t)s * GROWS * RS + rr * RS;
            if (lane & 1) {
                const double* hp = t == 0 ? h0 + ((int64_t)dir * n + row) * GH : out + (row * L + lp) * (2 * GH) + dir * GH;
                g_bulk_g2s(g_smem_u32(dst + 4 * GH), hp, GH * 8, bars + 8 * s);
            } else {
                g_bulk_g2s(g_smem_u32(dst), gates + ((row * L + l) * 2 + dir) * (4 * GH), 4 * GH * 8, bars + 8 * s);
            }
        }
    };
    issue(0);
    double dh[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dh[mt][0] = dh[mt][1] = 0.0;
    double2 dOn[MT];
    {
        const int l0 = dir ? 0 : L - 1;            // position of backward step 0
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r = 8 * mt + gid;
            const double* src = dout_last ? dout + (int64_t)(row0 + r) * (2 * GH) + dir * GH + j0
                                          : dout + ((int64_t)(row0 + r) * L + l0) * (2 * GH) + dir * GH + j0;
            dOn[mt] = (jok && r < nrows && !(dout_last && l0 != L - 1)) ? *reinterpret_cast<const double2*>(src) : make_double2(0.0, 0.0);
        }
    }

    for (int tb = 0; tb < L; ++tb) {
        const int t = L - 1 - tb;
        const int l = dir ? L - 1 - t : t;
        GRU_TRACE_B(0);
        if (tb + 1 < L) issue(tb + 1);                   // the other stage was drained before the last barrier
        GRU_TRACE_B(1);
        // dout of the NEXT step is requested now and consumed a whole step later: a global load issued and used inside the
        // same step exposed ~1.5 k cycles of DRAM latency in front of the gate arithmetic (profiles/r01/trace_gru_v4.txt)
        double2 dO[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dO[mt] = dOn[mt];
        if (tb + 1 < L) {
            const int ln = dir ? l + 1 : l - 1;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r = 8 * mt + gid;
                if (dout_last) {
                    const bool hit = jok && r < nrows && ln == L - 1;
                    dOn[mt] = hit ? *reinterpret_cast<const double2*>(dout + (int64_t)(row0 + r) * (2 * GH) + dir * GH + j0)
                                  : make_double2(0.0, 0.0);
                } else {
                    dOn[mt] = (jok && r < nrows) ? *reinterpret_cast<const double2*>(dout + ((int64_t)(row0 + r) * L + ln) * (2 * GH) + dir * GH + j0)
                                                 : make_double2(0.0, 0.0);
                }
            }
        }
        const int s = tb % GB_STAGES;
        GRU_TRACE_B(2);
        g_mbar_wait(bars + 8 * s, (uint32_t)(tb / GB_STAGES) & 1u);
        GRU_TRACE_B(3);
        const double* gs = ring + (int64_t)s * GROWS * RS;
        if (jok) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r = 8 * mt + gid;
                const double2 rr = *reinterpret_cast<const double2*>(gs + r * RS + j0);
                const double2 zz = *reinterpret_cast<const double2*>(gs + r * RS + GH + j0);
                const double2 nn = *reinterpret_cast<const double2*>(gs + r * RS + 2 * GH + j0);
                const double2 qq = *reinterpret_cast<const double2*>(gs + r * RS + 3 * GH + j0);
                const double2 hp = *reinterpret_cast<const double2*>(gs + r * RS + 4 * GH + j0);
                double dr[2], dz[2], dn[2], dq[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double rv = e ? rr.y : rr.x, zv = e ? zz.y : zz.x, nv = e ? nn.y : nn.x, qv = e ? qq.y : qq.x;
                    const double hv = e ? hp.y : hp.x;
                    const double d = (e ? dO[mt].y : dO[mt].x) + dh[mt][e];
                    dn[e] = d * (1.0 - zv) * (1.0 - nv * nv);
                    dz[e] = d * (hv - nv) * zv * (1.0 - zv);
                    dr[e] = dn[e] * qv * rv * (1.0 - rv);
                    dq[e] = dn[e] * rv;
                    dh[mt][e] = d * zv;
                }
                *reinterpret_cast<double2*>(dgh + r * GDS + j0) = make_double2(dr[0], dr[1]);
                *reinterpret_cast<double2*>(dgh + r * GDS + GH + j0) = make_double2(dz[0], dz[1]);
                *reinterpret_cast<double2*>(dgh + r * GDS + 2 * GH + j0) = make_double2(dq[0], dq[1]);
                if (r < nrows) {
                    double* gp = dgi + (((int64_t)(row0 + r) * L + l) * dgi_dirs + (dgi_dirs == 2 ? dir : 0)) * G3 + j0;
                    *reinterpret_cast<double2*>(gp) = make_double2(dr[0], dr[1]);
                    *reinterpret_cast<double2*>(gp + GH) = make_double2(dz[0], dz[1]);
                    *reinterpret_cast<double2*>(gp + 2 * GH) = make_double2(dn[0], dn[1]);
                }
            }
        }
        GRU_TRACE_B(4);
        __syncthreads();                           // dgh of this step complete; ring stage s drained
        GRU_TRACE_B(5);
        double acc[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 45; ++ks) {
            double a[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a[mt] = dgh[(8 * mt + gid) * GDS + 4 * ks + tig];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma_8x8x4(acc[mt][0], acc[mt][1], a[mt], B[ks]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            dh[mt][0] += acc[mt][0];
            dh[mt][1] += acc[mt][1];
        }
        GRU_TRACE_B(6);
        __syncthreads();                           // everyone done reading dgh before the next step overwrites it
        GRU_TRACE_B(7);
    }
    if (dh0 && jok) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r = 8 * mt + gid;
            if (r < nrows)
                *reinterpret_cast<double2*>(dh0 + ((int64_t)dir * n + row0 + r) * GH + j0) = make_double2(dh[mt][0], dh[mt][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Few rows (at most ~300 sequences per launch: small trees, leaf-sharded ranks): a CLUSTER of two CTAs shares the eight
// sequences of a row group.
//
// With one wave of 8-row CTAs the machine is mostly idle (32 CTAs per direction at 250 sequences) and a step of the
// recurrence costs what ONE SM needs for it: the recurrent product is 8 warps x 45 DMMAs = 1,440 cycles of that SM's fp64
// pipe however few rows there are (profiles/r02/trace_gru_rows8.txt), plus the gate arithmetic of 8 x 60 elements.  Here CTA
// c of the pair owns hidden units 32 c .. 32 c + 31 (four warps, one per SM sub-partition: 45 DMMAs each) and the matching
// half of the gate arithmetic.  What the other half needs - the new h (forward), dgh (backward) - goes into BOTH CTAs'
// shared memory: a plain store at home and `st.async` into the partner, which completes bytes on the partner's mbarrier.
// No fence and no barrier.cluster inside the loop (one per step costs 1.3-1.6 k cycles here: its release is a MEMBAR.GPU
// that waits for the step's global stores, profiles/r02/trace_gru_cluster_v1.txt).  Per operand buffer one mbarrier
// collects the four local warps' arrivals and one the partner's bytes; the product multiplies the columns written at home
// first and the partner's - ~450 cycles of DSMEM + mbarrier latency away - second, by which time they have arrived.
// The operand buffers are double-buffered and nothing else is needed: whoever writes buffer b for step t + 2 has consumed
// everything the readers of step t produced AFTER reading it.  A fifth warp per CTA issues the TMA copies, up to two
// steps ahead (full / empty mbarriers).  Per step at 250 rows: 2.30 k cycles forward, 2.07 k backward, against 3.24 k /
// 2.93 k for one 8-row CTA (profiles/r02/trace_gru_cluster_v3.txt, trace_gru_rows8.txt).  Splitting the reduction
// dimension over a second warp per sub-partition does not help: the fp64 pipe of a sub-partition takes a DMMA every
// ~17 cycles whoever issues it, and the extra exchange costs more than the single warp's 22 (trace_gru_cluster_v4_8warps.txt).
// Same arguments and layouts as gru_fwd2_kernel / gru_bwd_kernel; rows per cluster: 8.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GC_CL = 2;                      // CTAs per cluster
constexpr int GC_NW = 8 / GC_CL;              // consumer warps per CTA (one group of 8 hidden units each)
constexpr int GC_THREADS = 32 * (GC_NW + 1);  // + the producer warp
constexpr int GC_ROWS = 8;
constexpr int GC_UNITS1 = GH - 8 * GC_NW;     // hidden units of CTA 1 (28; CTA 0 has 32)

__device__ __forceinline__ uint32_t g_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t g_mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// 16 bytes into the partner's shared memory; completes 16 bytes of the transaction count of the partner's mbarrier
__device__ __forceinline__ void g_st_async(uint32_t addr, double a, double b, uint32_t bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(a),
                 "d"(b), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void g_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__host__ __device__ constexpr int gfc_smem(bool uv)
{
    // h double buffer | W^T | ring (UV: one V row per stage, + the U rows) | barriers | exponent table
    return 2 * GC_ROWS * GHS * 8 + (GH * GWS + 8) * 8 + GF2_STAGES * (uv ? 1 : GC_ROWS) * GRS * 8 + (uv ? GC_ROWS * GRS * 8 : 0) + 64 + 256;
}

template <bool UV>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_fwdc_kernel(const double* __restrict__ gi, const double* __restrict__ vrow,
                                                                 const double* __restrict__ whh, const double* __restrict__ bhh,
                                                                 const double* __restrict__ h0, int n, int L,
                                                                 double* __restrict__ out, double* __restrict__ gates, int gi_dirs)
{
    constexpr int STAGE_ROWS = UV ? 1 : GC_ROWS;
    extern __shared__ __align__(16) uint8_t gsm[];
    double* hbuf = reinterpret_cast<double*>(gsm);                                   // [2][8][68]
    double* wt = hbuf + 2 * GC_ROWS * GHS;                                            // [60][180]: wt[k][c] = W_hh[c][k]
    double* ring = wt + GH * GWS + 8;                                                 // [stage][8 or 1][184]
    double* ubuf = ring + GF2_STAGES * STAGE_ROWS * GRS;                              // [8][184] (UV only)
    double* etab = reinterpret_cast<double*>(gsm + gfc_smem(UV) - 256);
    const uint32_t full = g_smem_u32(gsm + gfc_smem(UV) - 320);                       // [stages]  ring stage landed
    const uint32_t empty = full + 8 * GF2_STAGES;                                     // [stages]  ring stage read by the 4 warps
    const uint32_t hloc = empty + 8 * GF2_STAGES;                                     // [2]       h buffer: this CTA's half written
    const uint32_t hrem = hloc + 16;                                                  // [2]       h buffer: the partner's half landed
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const uint32_t crank = g_cluster_ctarank();
    const bool producer = warp == GC_NW;
    const int wg = (int)crank * GC_NW + warp;                                         // hidden-unit group (consumers)
    const int dir = blockIdx.y;
    const int row0 = (blockIdx.x / GC_CL) * GC_ROWS;
    const int nrows = min(GC_ROWS, n - row0);
    const int j0 = wg * 8 + 2 * tig;
    const bool jok = !producer && j0 < GH;
    const int jb = wg * 8 + gid;
    const double* W = whh + (int64_t)dir * G3 * GH;
    // bytes of h the partner delivers per step: all 8 rows of its hidden units
    const uint32_t rx_bytes = (uint32_t)GC_ROWS * 8u * (crank == 0 ? (uint32_t)GC_UNITS1 : 8u * GC_NW);

    for (int idx = tid; idx < GH * GWS + 8; idx += GC_THREADS) {
        const int k = idx / GWS, c = idx - k * GWS;
        wt[idx] = k < GH ? W[(int64_t)c * GH + k] : 0.0;
    }
    if (tid < 32) etab[tid] = DRP_EXP2_TAB[tid];
    double bh[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bh[g][0] = jok ? bhh[dir * G3 + g * GH + j0] : 0.0;
        bh[g][1] = jok ? bhh[dir * G3 + g * GH + j0 + 1] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < GF2_STAGES; ++s) {
            g_mbar_init(full + 8 * s, 1);
            g_mbar_init(empty + 8 * s, GC_NW);
        }
        for (int b = 0; b < 2; ++b) {
            g_mbar_init(hloc + 8 * b, GC_NW);
            g_mbar_init(hrem + 8 * b, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int idx = tid; idx < GC_ROWS * GH; idx += GC_THREADS) {     // every CTA starts from all 60 units of h0
        const int r = idx / GH, j = idx - r * GH;
        hbuf[r * GHS + j] = r < nrows ? h0[((int64_t)dir * n + row0 + r) * GH + j] : 0.0;
    }
    if (UV) {
        for (int idx = tid; idx < GC_ROWS * G3; idx += GC_THREADS) {
            const int r = idx / G3, c = idx - r * G3;
            ubuf[r * GRS + c] = r < nrows ? gi[((int64_t)(row0 + r) * 2 + dir) * G3 + c] : 0.0;
        }
    }
    __syncthreads();
    g_cluster_sync();              // the partner is running, its barriers are initialised: remote stores may start

    if (producer) {                // the gi rows (UV: the V row) of step t into ring stage t % 2
        for (int t = 0; t < L; ++t) {
            const int s = t % GF2_STAGES;
            const int l = dir ? L - 1 - t : t;
            if (t >= GF2_STAGES) g_mbar_wait(empty + 8 * s, (uint32_t)(t / GF2_STAGES - 1) & 1u);
            if (UV) {
                if (lane == 0) {
                    g_mbar_expect_tx(full + 8 * s, G3 * 8);
                    g_bulk_g2s(g_smem_u32(ring + s * GRS), vrow + ((int64_t)l * 2 + dir) * G3, G3 * 8, full + 8 * s);
                }
            } else {
                if (lane == 0) g_mbar_expect_tx(full + 8 * s, (uint32_t)nrows * G3 * 8);
                __syncwarp();
                if (lane < nrows)
                    g_bulk_g2s(g_smem_u32(ring + (int64_t)s * GC_ROWS * GRS + lane * GRS),
                               gi + (((int64_t)(row0 + lane) * L + l) * gi_dirs + (gi_dirs == 2 ? dir : 0)) * G3, G3 * 8, full + 8 * s);
            }
        }
        g_cluster_sync();
        return;
    }
    const uint32_t hb_remote = g_mapa(g_smem_u32(hbuf), crank ^ 1u);
    const uint32_t hrem_remote = g_mapa(hrem, crank ^ 1u);
    double hold[2];
    hold[0] = jok ? hbuf[gid * GHS + j0] : 0.0;
    hold[1] = jok ? hbuf[gid * GHS + j0 + 1] : 0.0;

    // CR = rank in the pair, as a compile-time constant: the k-steps of the recurrent product that read this CTA's OWN half
    // of h (written by its four warps: a local barrier away) go first, the partner's half - ~450 cycles of DSMEM + mbarrier
    // latency away - second, by which time it has normally arrived.  Six accumulation chains (gate x k parity).
    auto run = [&](auto cr_tag) {
        constexpr int CR = decltype(cr_tag)::value;
        constexpr int LK0 = CR ? 8 : 0, LKN = CR ? 7 : 8, RK0 = CR ? 0 : 8, RKN = 15 - LKN;
        for (int t = 0; t < L; ++t) {
            const int l = dir ? L - 1 - t : t;
            const double* hc = hbuf + (t & 1) * GC_ROWS * GHS;
            const int nb = (t + 1) & 1;
            const int hn_off = nb * GC_ROWS * GHS + gid * GHS + j0;
            const uint32_t hpar = (uint32_t)((t - 1) >> 1) & 1u;
            GRU_TRACE(0);
            if (t > 0) g_mbar_wait(hloc + 8 * (t & 1), hpar);                  // h(t-1), this CTA's half
            GRU_TRACE(1);
            double acc[2][3][2];
#pragma unroll
            for (int g = 0; g < 3; ++g) acc[0][g][0] = bh[g][0], acc[0][g][1] = bh[g][1], acc[1][g][0] = acc[1][g][1] = 0.0;
#pragma unroll
            for (int i = 0; i < LKN; ++i) {
                const int ks = LK0 + i;
                const double a = hc[gid * GHS + 4 * ks + tig];
                double b[3];
#pragma unroll
                for (int g = 0; g < 3; ++g) b[g] = wt[(4 * ks + tig) * GWS + g * GH + jb];
#pragma unroll
                for (int g = 0; g < 3; ++g) dmma_8x8x4(acc[i & 1][g][0], acc[i & 1][g][1], a, b[g]);
            }
            GRU_TRACE(7);
            if (t > 0) g_mbar_wait(hrem + 8 * (t & 1), hpar);                  // h(t-1), the partner's half
#pragma unroll
            for (int i = 0; i < RKN; ++i) {
                const int ks = RK0 + i;
                const double a = hc[gid * GHS + 4 * ks + tig];
                double b[3];
#pragma unroll
                for (int g = 0; g < 3; ++g) b[g] = wt[(4 * ks + tig) * GWS + g * GH + jb];
#pragma unroll
                for (int g = 0; g < 3; ++g) dmma_8x8x4(acc[(LKN + i) & 1][g][0], acc[(LKN + i) & 1][g][1], a, b[g]);
            }
            const int s = t % GF2_STAGES;
            GRU_TRACE(2);
            g_mbar_wait(full + 8 * s, (uint32_t)(t / GF2_STAGES) & 1u);
            GRU_TRACE(3);
            const double* gs = ring + (int64_t)s * STAGE_ROWS * GRS;
            double hv[2] = {0.0, 0.0}, rv[2], zv[2], nv[2], qv[2];
            if (jok) {
                double xr[2], xz[2], xn[2];
                double2 ar, az, an;
                if (UV) {
                    const double* up = ubuf + gid * GRS;
                    const double2 ur = *reinterpret_cast<const double2*>(up + j0);
                    const double2 uz = *reinterpret_cast<const double2*>(up + GH + j0);
                    const double2 un = *reinterpret_cast<const double2*>(up + 2 * GH + j0);
                    const double2 vr = *reinterpret_cast<const double2*>(gs + j0);
                    const double2 vz = *reinterpret_cast<const double2*>(gs + GH + j0);
                    const double2 vn = *reinterpret_cast<const double2*>(gs + 2 * GH + j0);
                    ar = make_double2(ur.x + vr.x, ur.y + vr.y);
                    az = make_double2(uz.x + vz.x, uz.y + vz.y);
                    an = make_double2(un.x + vn.x, un.y + vn.y);
                } else {
                    ar = *reinterpret_cast<const double2*>(gs + gid * GRS + j0);
                    az = *reinterpret_cast<const double2*>(gs + gid * GRS + GH + j0);
                    an = *reinterpret_cast<const double2*>(gs + gid * GRS + 2 * GH + j0);
                }
                xr[0] = ar.x, xr[1] = ar.y, xz[0] = az.x, xz[1] = az.y, xn[0] = an.x, xn[1] = an.y;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    qv[e] = acc[0][2][e] + acc[1][2][e];                 // the accumulators started from b_hh
                    xr[e] += acc[0][0][e] + acc[1][0][e];
                    xz[e] += acc[0][1][e] + acc[1][1][e];
                    rv[e] = gru_sigmoid(xr[e], etab);
                    zv[e] = gru_sigmoid(xz[e], etab);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double gn = fma(rv[e], qv[e], xn[e]);
                    nv[e] = gru_tanh(gn, etab);
                    hv[e] = fma(zv[e], hold[e] - nv[e], nv[e]);
                    hv[e] = fma(0.0, (xr[e] + xz[e]) + gn, hv[e]);   // a NaN pre-activation reaches h (see gru_fwd2_kernel)
                    hold[e] = hv[e];
                }
                if (t + 1 < L) {                                     // (nobody reads the last h from shared memory)
                    *reinterpret_cast<double2*>(hbuf + hn_off) = make_double2(hv[0], hv[1]);
                    g_st_async(hb_remote + (uint32_t)hn_off * 8u, hv[0], hv[1], hrem_remote + 8u * nb);
                }
            }
            GRU_TRACE(4);
            __syncwarp();
            if (lane == 0) {
                if (t + 1 < L) {
                    g_mbar_arrive(hloc + 8 * nb);
                    if (warp == 0) g_mbar_expect_tx(hrem + 8 * nb, rx_bytes);  // the only arrival: announces the partner's bytes
                }
                g_mbar_arrive(empty + 8 * s);                                   // this warp has read ring stage s
            }
            GRU_TRACE(6);
            if (jok && gid < nrows) {
                const int64_t site = (int64_t)(row0 + gid) * L + l;
                *reinterpret_cast<double2*>(out + site * (2 * GH) + dir * GH + j0) = make_double2(hv[0], hv[1]);
                if (gates) {
                    double* gp = gates + (site * 2 + dir) * (4 * GH) + j0;
                    *reinterpret_cast<double2*>(gp) = make_double2(rv[0], rv[1]);
                    *reinterpret_cast<double2*>(gp + GH) = make_double2(zv[0], zv[1]);
                    *reinterpret_cast<double2*>(gp + 2 * GH) = make_double2(nv[0], nv[1]);
                    *reinterpret_cast<double2*>(gp + 3 * GH) = make_double2(qv[0], qv[1]);
                }
            }
            GRU_TRACE(5);
        }
    };
    if (crank == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 1>{});
    g_cluster_sync();              // nobody leaves while its partner could still be addressed
}

__host__ __device__ constexpr int gbc_smem() { return 2 * GC_ROWS * GDS * 8 + GB_STAGES * GC_ROWS * GB_ROW_STRIDE * 8 + 64; }

__global__ void __launch_bounds__(GC_THREADS, 1) gru_bwdc_kernel(const double* __restrict__ dout, const double* __restrict__ out,
                                                                 const double* __restrict__ gates, const double* __restrict__ whh,
                                                                 const double* __restrict__ h0, int n, int L,
                                                                 double* __restrict__ dgi, double* __restrict__ dh0, int dgi_dirs,
                                                                 int dout_last)
{
    constexpr int RS = GB_ROW_STRIDE;
    extern __shared__ __align__(16) uint8_t gsm[];
    double* dgh = reinterpret_cast<double*>(gsm);                                    // [2][8][180]
    double* ring = dgh + 2 * GC_ROWS * GDS;                                           // [stages][8][312]
    const uint32_t full = g_smem_u32(gsm + gbc_smem() - 64);
    const uint32_t empty = full + 8 * GB_STAGES;
    const uint32_t dloc = empty + 8 * GB_STAGES;                                      // [2] dgh buffer: this CTA's columns written
    const uint32_t drem = dloc + 16;                                                  // [2] dgh buffer: the partner's columns landed
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const uint32_t crank = g_cluster_ctarank();
    const bool producer = warp == GC_NW;
    const int wg = (int)crank * GC_NW + warp;
    const int dir = blockIdx.y;
    const int row0 = (blockIdx.x / GC_CL) * GC_ROWS;
    const int nrows = min(GC_ROWS, n - row0);
    const int j0 = wg * 8 + 2 * tig;
    const bool jok = !producer && j0 < GH;
    const int jb = wg * 8 + gid;
    const double* W = whh + (int64_t)dir * G3 * GH;
    const uint32_t rx_bytes = 3u * GC_ROWS * 8u * (crank == 0 ? (uint32_t)GC_UNITS1 : 8u * GC_NW);

    if (tid == 0) {
        for (int s = 0; s < GB_STAGES; ++s) {
            g_mbar_init(full + 8 * s, 1);
            g_mbar_init(empty + 8 * s, GC_NW);
        }
        for (int b = 0; b < 2; ++b) {
            g_mbar_init(dloc + 8 * b, GC_NW);
            g_mbar_init(drem + 8 * b, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    g_cluster_sync();

    // backward step tb = 0..L-1 handles forward step t = L-1-tb, i.e. position l = dir ? tb : L-1-tb
    if (producer) {                // saved gates + h_prev of the 8 rows, one copy per lane
        for (int tb = 0; tb < L; ++tb) {
            const int s = tb % GB_STAGES;
            const int t = L - 1 - tb;
            const int l = dir ? L - 1 - t : t;
            const int lp = dir ? l + 1 : l - 1;       // position of h_prev (outside [0,L): h0)
            if (tb >= GB_STAGES) g_mbar_wait(empty + 8 * s, (uint32_t)(tb / GB_STAGES - 1) & 1u);
            if (lane == 0) g_mbar_expect_tx(full + 8 * s, (uint32_t)nrows * GB_ROW_BYTES);
            __syncwarp();
            const int rr = lane >> 1;
            if (lane < 2 * GC_ROWS && rr < nrows) {
                const int64_t row = row0 + rr;
                double* dst = ring + (int64_t)s * GC_ROWS * RS + rr * RS;
                if (lane & 1) {
                    const double* hp = t == 0 ? h0 + ((int64_t)dir * n + row) * GH : out + (row * L + lp) * (2 * GH) + dir * GH;
                    g_bulk_g2s(g_smem_u32(dst + 4 * GH), hp, GH * 8, full + 8 * s);
                } else {
                    g_bulk_g2s(g_smem_u32(dst), gates + ((row * L + l) * 2 + dir) * (4 * GH), 4 * GH * 8, full + 8 * s);
                }
            }
        }
        g_cluster_sync();
        return;
    }
    double B[45];                                 // B[k][n] = W_hh[c = 4 ks + tig][j = jb]
#pragma unroll
    for (int ks = 0; ks < 45; ++ks) B[ks] = jb < GH ? W[(int64_t)(4 * ks + tig) * GH + jb] : 0.0;
    const uint32_t dg_remote = g_mapa(g_smem_u32(dgh), crank ^ 1u);
    const uint32_t drem_remote = g_mapa(drem, crank ^ 1u);
    double dh[2] = {0.0, 0.0};
    const bool rok = jok && gid < nrows;
    double2 dOn;
    {
        const int l0 = dir ? 0 : L - 1;            // position of backward step 0
        const double* src = dout_last ? dout + (int64_t)(row0 + gid) * (2 * GH) + dir * GH + j0
                                      : dout + ((int64_t)(row0 + gid) * L + l0) * (2 * GH) + dir * GH + j0;
        dOn = (rok && !(dout_last && l0 != L - 1)) ? *reinterpret_cast<const double2*>(src) : make_double2(0.0, 0.0);
    }

    // CR = rank in the pair at compile time (W_hh sits in registers: static indices): the columns of dgh this CTA's own
    // warps wrote are multiplied first, the partner's - still in flight - second; six accumulation chains
    auto run = [&](auto cr_tag) {
        constexpr int CR = decltype(cr_tag)::value;
        constexpr int LK0 = CR ? 8 : 0, LKN = CR ? 7 : 8, RK0 = CR ? 0 : 8, RKN = 15 - LKN;
        for (int tb = 0; tb < L; ++tb) {
            const int t = L - 1 - tb;
            const int l = dir ? L - 1 - t : t;
            GRU_TRACE_B(0);
            const double2 dO = dOn;                    // requested a whole step ago (see gru_bwd_kernel)
            if (tb + 1 < L) {
                const int ln = dir ? l + 1 : l - 1;
                if (dout_last) {
                    dOn = (rok && ln == L - 1) ? *reinterpret_cast<const double2*>(dout + (int64_t)(row0 + gid) * (2 * GH) + dir * GH + j0)
                                               : make_double2(0.0, 0.0);
                } else {
                    dOn = rok ? *reinterpret_cast<const double2*>(dout + ((int64_t)(row0 + gid) * L + ln) * (2 * GH) + dir * GH + j0)
                              : make_double2(0.0, 0.0);
                }
            }
            const int s = tb % GB_STAGES, b = tb & 1;
            const int dg_off = b * GC_ROWS * GDS + gid * GDS + j0;
            const uint32_t dpar = (uint32_t)(tb >> 1) & 1u;
            GRU_TRACE_B(2);
            g_mbar_wait(full + 8 * s, (uint32_t)(tb / GB_STAGES) & 1u);
            GRU_TRACE_B(3);
            const double* gs = ring + (int64_t)s * GC_ROWS * RS + gid * RS;
            double dr[2], dz[2], dn[2], dq[2];
            if (jok) {
                const double2 rr = *reinterpret_cast<const double2*>(gs + j0);
                const double2 zz = *reinterpret_cast<const double2*>(gs + GH + j0);
                const double2 nn = *reinterpret_cast<const double2*>(gs + 2 * GH + j0);
                const double2 qq = *reinterpret_cast<const double2*>(gs + 3 * GH + j0);
                const double2 hp = *reinterpret_cast<const double2*>(gs + 4 * GH + j0);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double rv = e ? rr.y : rr.x, zv = e ? zz.y : zz.x, nv = e ? nn.y : nn.x, qv = e ? qq.y : qq.x;
                    const double hv = e ? hp.y : hp.x;
                    const double d = (e ? dO.y : dO.x) + dh[e];
                    dn[e] = d * (1.0 - zv) * (1.0 - nv * nv);
                    dz[e] = d * (hv - nv) * zv * (1.0 - zv);
                    dr[e] = dn[e] * qv * rv * (1.0 - rv);
                    dq[e] = dn[e] * rv;
                    dh[e] = d * zv;
                }
                *reinterpret_cast<double2*>(dgh + dg_off) = make_double2(dr[0], dr[1]);
                *reinterpret_cast<double2*>(dgh + dg_off + GH) = make_double2(dz[0], dz[1]);
                *reinterpret_cast<double2*>(dgh + dg_off + 2 * GH) = make_double2(dq[0], dq[1]);
                g_st_async(dg_remote + (uint32_t)dg_off * 8u, dr[0], dr[1], drem_remote + 8u * b);
                g_st_async(dg_remote + (uint32_t)(dg_off + GH) * 8u, dz[0], dz[1], drem_remote + 8u * b);
                g_st_async(dg_remote + (uint32_t)(dg_off + 2 * GH) * 8u, dq[0], dq[1], drem_remote + 8u * b);
            }
            GRU_TRACE_B(4);
            __syncwarp();
            if (lane == 0) {
                g_mbar_arrive(dloc + 8 * b);
                if (warp == 0) g_mbar_expect_tx(drem + 8 * b, rx_bytes);       // the only arrival: announces the partner's bytes
                g_mbar_arrive(empty + 8 * s);                                  // this warp has read ring stage s
            }
            GRU_TRACE_B(1);
            if (rok) {
                double* gp = dgi + (((int64_t)(row0 + gid) * L + l) * dgi_dirs + (dgi_dirs == 2 ? dir : 0)) * G3 + j0;
                *reinterpret_cast<double2*>(gp) = make_double2(dr[0], dr[1]);
                *reinterpret_cast<double2*>(gp + GH) = make_double2(dz[0], dz[1]);
                *reinterpret_cast<double2*>(gp + 2 * GH) = make_double2(dn[0], dn[1]);
            }
            GRU_TRACE_B(7);
            // dh_prev[:, own units] = dh z + dgh W_hh[:, own units]
            const double* dg = dgh + b * GC_ROWS * GDS + gid * GDS + tig;
            double acc[6][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
            g_mbar_wait(dloc + 8 * b, dpar);                                   // the columns this CTA's warps wrote
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int i = 0; i < LKN; ++i) {
                    const int ks = g * 15 + LK0 + i, c = (g * LKN + i) % 6;
                    dmma_8x8x4(acc[c][0], acc[c][1], dg[4 * ks], B[ks]);
                }
            GRU_TRACE_B(5);
            g_mbar_wait(drem + 8 * b, dpar);                                   // the partner's columns
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int i = 0; i < RKN; ++i) {
                    const int ks = g * 15 + RK0 + i, c = (3 * LKN + g * RKN + i) % 6;
                    dmma_8x8x4(acc[c][0], acc[c][1], dg[4 * ks], B[ks]);
                }
            dh[0] += ((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0])) + (acc[4][0] + acc[5][0]);
            dh[1] += ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1])) + (acc[4][1] + acc[5][1]);
            GRU_TRACE_B(6);
        }
    };
    if (crank == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 1>{});
    if (dh0 && rok) *reinterpret_cast<double2*>(dh0 + ((int64_t)dir * n + row0 + gid) * GH + j0) = make_double2(dh[0], dh[1]);
    g_cluster_sync();              // nobody leaves while its partner could still be addressed
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradients of the recurrence: dW_hh[dir] = sum_p dgh[p]^T h_prev[p], db_hh[dir] = sum_p dgh[p] over all
// (sequence, step) pairs p, with dgh = dgi except that the n-part is multiplied by r.  A [180 x 60] += [180 x K][K x 60]
// product with K = n L, split over the grid along K; every CTA keeps its partial product in registers (fp64 mma.sync
// accumulators), operand rows arrive by TMA bulk copies, partials are summed by a second, deterministic kernel.
// The bias gradient rides along as column 60 of the padded h_prev tile (a column of ones).
//   part [n_cta/2][2][184][64]
// ---------------------------------------------------------------------------------------------------------------
constexpr int GW_PAIRS = 32;                                   // pairs per pipeline stage (8 k-steps)
constexpr int GW_ROW = (G3 + GH + GH);                          // doubles per staged pair: dgi | r | h_prev
constexpr int GW_STAGE_BYTES = GW_PAIRS * GW_ROW * 8;           // 76800
constexpr int GW_STAGES = 2;
constexpr int GW_SMEM = GW_STAGES * GW_STAGE_BYTES + 64;
constexpr int GW_MT = 23;                                      // 184 = 23 m-tiles cover the 180 gate columns

constexpr int GW_THREADS = GTHREADS + 64;      // 8 consumer warps (one 8-column group of [h_prev | 1] each) + 2 producer warps

// Warps 8 and 9 only issue the bulk copies (96 per 32-pair stage, ~60 cycles of issue each: on the consumer warps they
// cost 1.4 k of the 8.9 k cycles of a stage, profiles/r01/trace_gru_v6.txt); the consumers never leave the DMMA loop:
// full[s] (transaction barrier) says the stage has landed, empty[s] (one arrival per consumer warp) that it may be refilled.
__global__ void __launch_bounds__(GW_THREADS, 1) gru_wgrad_kernel(const double* __restrict__ dgi, const double* __restrict__ gates,
                                                                  const double* __restrict__ out, const double* __restrict__ h0,
                                                                  int n, int L, double* __restrict__ part, int ndir,
                                                                  int dgi_dirs)
{
    extern __shared__ __align__(16) uint8_t gsm[];
    double* ring = reinterpret_cast<double*>(gsm);
    const uint32_t full = g_smem_u32(gsm + GW_STAGES * GW_STAGE_BYTES);
    const uint32_t empty = full + 8 * GW_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int dir = ndir == 2 ? (blockIdx.x & 1) : 0;
    const int chunk = ndir == 2 ? (blockIdx.x >> 1) : blockIdx.x, nchunks = ndir == 2 ? (gridDim.x >> 1) : gridDim.x;
    const int64_t P = (int64_t)n * L;
    const int64_t nblk = (P + GW_PAIRS - 1) / GW_PAIRS;
    const int64_t b0 = nblk * chunk / nchunks, b1 = nblk * (chunk + 1) / nchunks;
    if (tid == 0) {
        for (int s = 0; s < GW_STAGES; ++s) {
            g_mbar_init(full + 8 * s, 1);
            g_mbar_init(empty + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp >= 8) {
        // ---------------- producers: 48 of the 96 copies of a stage each
        const int pw = warp - 8;
        for (int64_t blk = b0; blk < b1; ++blk) {
            const int it = (int)(blk - b0);
            const int s = it % GW_STAGES;
            if (it >= GW_STAGES) g_mbar_wait(empty + 8 * s, (uint32_t)(it / GW_STAGES - 1) & 1u);
            const int valid = (int)min((int64_t)GW_PAIRS, P - blk * GW_PAIRS);
            // the first producer's arrival carries the byte count, so the phase cannot complete before it
            if (pw == 0 && lane == 0) g_mbar_expect_tx(full + 8 * s, (uint32_t)valid * GW_ROW * 8);
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int task = pw * 48 + rep * 32 + lane;              // 0..95 = pair * 3 + which
                if (rep == 1 && lane >= 16) break;
                const int pr = task / 3, which = task - 3 * pr;
                if (pr >= valid) continue;
                const int64_t p = blk * GW_PAIRS + pr;
                double* dst = ring + (int64_t)s * GW_PAIRS * GW_ROW + pr * GW_ROW;
                if (which == 0) {
                    g_bulk_g2s(g_smem_u32(dst), dgi + (p * dgi_dirs + (dgi_dirs == 2 ? dir : 0)) * G3, G3 * 8, full + 8 * s);
                } else if (which == 1) {
                    g_bulk_g2s(g_smem_u32(dst + G3), gates + (p * 2 + dir) * (4 * GH), GH * 8, full + 8 * s);
                } else {
                    const int64_t i = p / L;
                    const int l = (int)(p - i * L);
                    const bool first = dir ? (l == L - 1) : (l == 0);  // first step of the direction: h_prev = h0
                    const double* hp = first ? h0 + ((int64_t)dir * n + i) * GH : out + (dir ? p + 1 : p - 1) * (2 * GH) + dir * GH;
                    g_bulk_g2s(g_smem_u32(dst + G3 + GH), hp, GH * 8, full + 8 * s);
                }
            }
        }
        return;
    }
    // ---------------- consumers
    double acc[GW_MT][2];
#pragma unroll
    for (int mt = 0; mt < GW_MT; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
    const int jb = warp * 8 + gid;                 // column of h_prev (60: ones -> bias gradient; 61..63: zero)
    for (int64_t blk = b0; blk < b1; ++blk) {
        const int it = (int)(blk - b0);
        const int s = it % GW_STAGES;
        GRU_TRACE_W(0);
        g_mbar_wait(full + 8 * s, (uint32_t)(it / GW_STAGES) & 1u);
        GRU_TRACE_W(1);
        double* st = ring + (int64_t)s * GW_PAIRS * GW_ROW;
        const int valid = (int)min((int64_t)GW_PAIRS, P - blk * GW_PAIRS);
        if (valid < GW_PAIRS) {                    // ragged last block: rows that were not loaded must not contribute
            for (int idx = tid; idx < (GW_PAIRS - valid) * GW_ROW; idx += GTHREADS) st[valid * GW_ROW + idx] = 0.0;
            g_named_barrier(1, GTHREADS);
        }
        GRU_TRACE_W(2);
#pragma unroll 2
        for (int ks = 0; ks < GW_PAIRS / 4; ++ks) {
            const double* row = st + (4 * ks + tig) * GW_ROW;
            const double b = jb < GH ? row[G3 + GH + jb] : (jb == GH ? 1.0 : 0.0);
#pragma unroll
            for (int mt = 0; mt < GW_MT; ++mt) {
                const int c = 8 * mt + gid;
                double a = c < G3 ? row[c] : 0.0;
                if (mt >= 15) a *= (c < G3 ? row[G3 + c - 2 * GH] : 0.0);       // n-part: times r
                dmma_8x8x4(acc[mt][0], acc[mt][1], a, b);
            }
        }
        GRU_TRACE_W(3);
        __syncwarp();
        if (lane == 0) g_mbar_arrive(empty + 8 * s);   // this warp is done with stage s
        GRU_TRACE_W(4);
    }
    double* dst = part + ((int64_t)chunk * 2 + dir) * (GW_MT * 8) * 64;
#pragma unroll
    for (int mt = 0; mt < GW_MT; ++mt) {
        const int c = 8 * mt + gid, j = warp * 8 + 2 * tig;
        *reinterpret_cast<double2*>(dst + (int64_t)c * 64 + j) = make_double2(acc[mt][0], acc[mt][1]);
    }
}

// dwhh [2][180][60], dbhh [2][180] = sum over chunks of part (fixed order: deterministic)
__global__ void __launch_bounds__(256) gru_wgrad_reduce_kernel(const double* __restrict__ part, int nchunks,
                                                               double* __restrict__ dwhh, double* __restrict__ dbhh, int ndir)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // over [ndir][180][61]
    if (idx >= ndir * G3 * (GH + 1)) return;
    const int j = idx % (GH + 1), c = (idx / (GH + 1)) % G3, dir = idx / ((GH + 1) * G3);
    double s = 0.0;
    for (int k = 0; k < nchunks; ++k) s += part[(((int64_t)k * 2 + dir) * (GW_MT * 8) + c) * 64 + j];
    if (j < GH) dwhh[((int64_t)dir * G3 + c) * GH + j] = s;
    else dbhh[dir * G3 + c] = s;
}

}  // namespace

// ------------------------------------------------- entry points -----------------------------------------------
DRP_API int drp_gru_hidden_size(void) { return GH; }

// rows per CTA = 8 MT: the largest of 32 / 16 / 8 for which one wave of CTAs (2 directions) still covers the batch
static int gru_pick_mt(int n, int ndir = 2)
{
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const char* e = getenv("DRP_GRU_ROWS");
    if (e) {
        const int r = atoi(e);
        if (r == 8 || r == 16 || r == 32) return r / 8;
    }
    if (ndir * ((n + 7) / 8) <= n_sm) return 1;
    if (ndir * ((n + 15) / 16) <= n_sm) return 2;
    return 4;
}

// Pairs of CTAs per 8-row group (gru_fwdc_kernel / gru_bwdc_kernel) whenever the pairs of ALL row groups and directions
// are resident at once; DRP_GRU_CLUSTER=0 switches the form off, =1 forces it for any size.
static bool gru_use_cluster(int n, int ndir)
{
    static int mode = -1, n_sm = 0;
    if (mode < 0) {
        const char* e = getenv("DRP_GRU_CLUSTER");
        mode = e ? (atoi(e) ? 1 : 0) : 2;
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    if (mode != 2) return mode == 1;
    if (getenv("DRP_GRU_ROWS")) return false;                       // an explicit rows-per-CTA request means the classic kernels
    return ndir * ((n + GC_ROWS - 1) / GC_ROWS) * GC_CL <= n_sm;
}

template <typename... KArgs, typename... Args>
static void gru_cluster_launch(void (*kern)(KArgs...), int n, int ndir, int smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((n + GC_ROWS - 1) / GC_ROWS) * GC_CL, ndir);
    cfg.blockDim = dim3(GC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GC_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);                        // an error stays in cudaGetLastError for drp_launch_status
}

template <bool UV>
static void gru_fwdc_launch(const double* gi, const double* v, const double* whh, const double* bhh, const double* h0, int n,
                            int L, double* out, double* gates, cudaStream_t st, int ndir)
{
    constexpr int smem = gfc_smem(UV);
    if (drp_first_on_device(60 + (int)UV)) {
        cudaFuncSetAttribute(gru_fwdc_kernel<UV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    DRP_LAUNCH(DRP_K_GRU_FWD, 2.0 * (double)n * L * ndir * G3 * GH, st);
    gru_cluster_launch(gru_fwdc_kernel<UV>, n, ndir, smem, st, gi, v, whh, bhh, h0, n, L, out, gates, ndir);
}

static void gru_bwdc_launch(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                            int n, int L, double* dgi, double* dh0, int ndir, int dgi_dirs, cudaStream_t st, int dout_last)
{
    constexpr int smem = gbc_smem();
    if (drp_first_on_device(62)) {
        cudaFuncSetAttribute(gru_bwdc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    DRP_LAUNCH(DRP_K_GRU_BWD, 2.0 * (double)n * L * ndir * G3 * GH, st);
    gru_cluster_launch(gru_bwdc_kernel, n, ndir, smem, st, dout, out, gates, whh, h0, n, L, dgi, dh0, dgi_dirs, dout_last);
}

static int gru_fwd_generation()
{
    static int gen = 0;
    if (!gen) {
        const char* e = getenv("DRP_GRU_FWD");
        gen = (e && atoi(e) == 1) ? 1 : 2;
    }
    return gen;
}

template <bool UV, int MT>
static void gru_fwd2_launch(const double* gi, const double* v, const double* whh, const double* bhh, const double* h0, int n,
                            int L, double* out, double* gates, cudaStream_t st, int ndir)
{
    constexpr int WS = MT >= 2 ? 2 : 1;
    constexpr int smem = gf2_smem(8 * MT, UV);
    if (drp_first_on_device(32 + 8 * (int)UV + MT)) {
        cudaFuncSetAttribute(gru_fwd2_kernel<UV, MT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    DRP_LAUNCH(DRP_K_GRU_FWD, 2.0 * (double)n * L * ndir * G3 * GH, st);
    static int stagger = -1;
    if (stagger < 0) {
        const char* e = getenv("DRP_GRU_HANDOFF");
        stagger = e ? atoi(e) : 0;                 // 0: the halves run free (they fall into lockstep: measured best); else: product token
    }
    gru_fwd2_kernel<UV, MT, WS><<<dim3((n + 8 * MT - 1) / (8 * MT), ndir), 256 * WS, smem, st>>>(gi, v, whh, bhh, h0, n, L, out,
                                                                                              gates, ndir, stagger);
}

template <bool UV, int MT>
static void gru_fwd_launch(const double* gi, const double* v, const double* whh, const double* bhh, const double* h0, int n,
                           int L, double* out, double* gates, cudaStream_t st, int ndir = 2)
{
    if (gru_fwd_generation() == 2) {
        gru_fwd2_launch<UV, MT>(gi, v, whh, bhh, h0, n, L, out, gates, st, ndir);
        return;
    }
    constexpr int smem = gf_smem(8 * MT, UV);
    if (drp_first_on_device(44 + 8 * (int)UV + MT)) {
        cudaFuncSetAttribute(gru_fwd_kernel<UV, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    DRP_LAUNCH(DRP_K_GRU_FWD, 2.0 * (double)n * L * ndir * G3 * GH, st);
    gru_fwd_kernel<UV, MT><<<dim3((n + 8 * MT - 1) / (8 * MT), ndir), GTHREADS, smem, st>>>(gi, v, whh, bhh, h0, n, L, out, gates,
                                                                                        ndir);
}

template <int MT>
static void gru_bwd_launch(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                           int n, int L, double* dgi, double* dh0, int ndir, int dgi_dirs, cudaStream_t st, int dout_last = 0)
{
    constexpr int smem = gb_smem(8 * MT);
    if (drp_first_on_device(24 + MT)) {
        cudaFuncSetAttribute(gru_bwd_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    DRP_LAUNCH(DRP_K_GRU_BWD, 2.0 * (double)n * L * ndir * G3 * GH, st);
    gru_bwd_kernel<MT><<<dim3((n + 8 * MT - 1) / (8 * MT), ndir), GTHREADS, smem, st>>>(dout, out, gates, whh, h0, n, L, dgi, dh0,
                                                                                    dgi_dirs, dout_last);
}

DRP_API int drp_gru_fwd_f64(const double* gi, const double* whh, const double* bhh, const double* h0, int n, int L, int H,
                            double* out, double* gates, void* stream)
{
    if (!gi || !whh || !bhh || !h0) return -1;
    if (n <= 0 || L <= 0) return -5;
    if (H != GH) return -7;
    if (!out) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    if (gru_use_cluster(n, 2)) {
        gru_fwdc_launch<false>(gi, nullptr, whh, bhh, h0, n, L, out, gates, st, 2);
        return drp_launch_status();
    }
    switch (gru_pick_mt(n)) {
        case 1: gru_fwd_launch<false, 1>(gi, nullptr, whh, bhh, h0, n, L, out, gates, st); break;
        case 2: gru_fwd_launch<false, 2>(gi, nullptr, whh, bhh, h0, n, L, out, gates, st); break;
        default: gru_fwd_launch<false, 4>(gi, nullptr, whh, bhh, h0, n, L, out, gates, st); break;
    }
    return drp_launch_status();
}

// Forward direction only: gi0 [n, L, 3H]; `out [n,L,2H]` / `gates [n,L,2,4,H]` keep the two-direction layout and only
// their direction-0 halves are written (the encoder's loss path, see ops._GRUEncoderLast).
DRP_API int drp_gru_fwd_dir0_f64(const double* gi0, const double* whh, const double* bhh, const double* h0, int n, int L,
                                 int H, double* out, double* gates, void* stream)
{
    if (!gi0 || !whh || !bhh || !h0) return -1;
    if (n <= 0 || L <= 0) return -5;
    if (H != GH) return -7;
    if (!out) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    if (gru_use_cluster(n, 1)) {
        gru_fwdc_launch<false>(gi0, nullptr, whh, bhh, h0, n, L, out, gates, st, 1);
        return drp_launch_status();
    }
    switch (gru_pick_mt(n, 1)) {
        case 1: gru_fwd_launch<false, 1>(gi0, nullptr, whh, bhh, h0, n, L, out, gates, st, 1); break;
        case 2: gru_fwd_launch<false, 2>(gi0, nullptr, whh, bhh, h0, n, L, out, gates, st, 1); break;
        default: gru_fwd_launch<false, 4>(gi0, nullptr, whh, bhh, h0, n, L, out, gates, st, 1); break;
    }
    return drp_launch_status();
}

// Same recurrence with gi[i][l] = U[i] + V[l]:  U [n,2,3H], V [L,2,3H]  (the decoder's tiled input, never materialised).
DRP_API int drp_gru_fwd_uv_f64(const double* U, const double* V, const double* whh, const double* bhh, const double* h0,
                               int n, int L, int H, double* out, double* gates, void* stream)
{
    if (!U || !V || !whh || !bhh || !h0) return -1;
    if (n <= 0 || L <= 0) return -6;
    if (H != GH) return -8;
    if (!out) return -9;
    cudaStream_t st = (cudaStream_t)stream;
    if (gru_use_cluster(n, 2)) {
        gru_fwdc_launch<true>(U, V, whh, bhh, h0, n, L, out, gates, st, 2);
        return drp_launch_status();
    }
    switch (gru_pick_mt(n)) {
        case 1: gru_fwd_launch<true, 1>(U, V, whh, bhh, h0, n, L, out, gates, st); break;
        case 2: gru_fwd_launch<true, 2>(U, V, whh, bhh, h0, n, L, out, gates, st); break;
        default: gru_fwd_launch<true, 4>(U, V, whh, bhh, h0, n, L, out, gates, st); break;
    }
    return drp_launch_status();
}

static int gru_bwd_dispatch(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                            int n, int L, int H, double* dgi, double* dh0, int ndir, int dgi_dirs, void* stream, int dout_last = 0)
{
    if (!dout || !out || !gates || !whh || !h0) return -1;
    if (n <= 0 || L <= 0) return -6;
    if (H != GH) return -8;
    if (!dgi) return -9;
    cudaStream_t st = (cudaStream_t)stream;
    if (gru_use_cluster(n, ndir)) {
        gru_bwdc_launch(dout, out, gates, whh, h0, n, L, dgi, dh0, ndir, dgi_dirs, st, dout_last);
        return drp_launch_status();
    }
    switch (gru_pick_mt(n, ndir)) {
        case 1: gru_bwd_launch<1>(dout, out, gates, whh, h0, n, L, dgi, dh0, ndir, dgi_dirs, st, dout_last); break;
        case 2: gru_bwd_launch<2>(dout, out, gates, whh, h0, n, L, dgi, dh0, ndir, dgi_dirs, st, dout_last); break;
        default: gru_bwd_launch<4>(dout, out, gates, whh, h0, n, L, dgi, dh0, ndir, dgi_dirs, st, dout_last); break;
    }
    return drp_launch_status();
}

DRP_API int drp_gru_bwd_f64(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                            int n, int L, int H, double* dgi, double* dh0, void* stream)
{
    return gru_bwd_dispatch(dout, out, gates, whh, h0, n, L, H, dgi, dh0, 2, 2, stream);
}

// Forward direction only (direction 0): dgi0 [n,L,3H] compact, dh0 [n,H] (first direction's slice of a [2,n,H] buffer is
// fine).  Used when the loss only sees the hidden state at the last position (the encoder, S/models_utils.py:57): the
// reverse direction then needs a single cell step, which the caller does with dense ops.
DRP_API int drp_gru_bwd_dir0_f64(const double* dout, const double* out, const double* gates, const double* whh,
                                 const double* h0, int n, int L, int H, double* dgi0, double* dh0, void* stream)
{
    return gru_bwd_dispatch(dout, out, gates, whh, h0, n, L, H, dgi0, dh0, 1, 1, stream);
}

// Same with the gradient given at the last position only: dlast [n, 2H] (its direction-0 half is read); the gradient of
// every other hidden state is zero.
DRP_API int drp_gru_bwd_last_dir0_f64(const double* dlast, const double* out, const double* gates, const double* whh,
                                      const double* h0, int n, int L, int H, double* dgi0, double* dh0, void* stream)
{
    return gru_bwd_dispatch(dlast, out, gates, whh, h0, n, L, H, dgi0, dh0, 1, 1, stream, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// Sums of dgi [n, L, C] over the positions (-> dU [n, C]) and, per block of rows, over the rows (-> dVpart [nblk, L, C]):
// the gradients of the decoder's broadcast input terms U[i] + V[l], in ONE pass over dgi (two torch reductions read the
// 2.9 GB tensor twice).  C even, C / 2 <= 256 threads; the caller adds the nblk partial dV blocks.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GS_ROWS = 16;                                      // rows per CTA at most
constexpr int GS_LU = 4;                                         // positions per iteration when a CTA has few rows
// grid (row blocks, position chunks).  With few rows per CTA (a leaf-sharded rank: 250 rows -> 2 per CTA) one position per
// iteration leaves two loads in flight per thread and the kernel runs at DRAM latency (0.47 ms for 360 MB); the positions are
// therefore also split over blockIdx.y (dU then comes out as per-chunk partials the caller adds) and walked four at a time.
__global__ void __launch_bounds__(256) gru_input_sums_kernel(const double* __restrict__ dgi, int n, int L, int C, int rows_per_blk,
                                                             int l_per_blk, double* __restrict__ dUpart, double* __restrict__ dVpart)
{
    const int c2 = threadIdx.x;
    if (2 * c2 >= C) return;
    const int r0 = blockIdx.x * rows_per_blk;
    const int nr = min(rows_per_blk, n - r0);
    const int l0 = blockIdx.y * l_per_blk, l1 = min(L, l0 + l_per_blk);
    double2 du[GS_ROWS];
#pragma unroll
    for (int r = 0; r < GS_ROWS; ++r) du[r] = make_double2(0.0, 0.0);
    const double2* base = reinterpret_cast<const double2*>(dgi) + (int64_t)r0 * L * (C / 2) + c2;
    double2* dv = reinterpret_cast<double2*>(dVpart) + (int64_t)blockIdx.x * L * (C / 2) + c2;
    if (nr <= GS_ROWS / GS_LU) {
        constexpr int RR = GS_ROWS / GS_LU;
        for (int l = l0; l < l1; l += GS_LU) {
            double2 v[GS_LU][RR];
#pragma unroll
            for (int u = 0; u < GS_LU; ++u)
#pragma unroll
                for (int r = 0; r < RR; ++r)
                    v[u][r] = (r < nr && l + u < l1) ? __ldcs(base + ((int64_t)r * L + l + u) * (C / 2)) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < GS_LU; ++u) {
                double2 s = make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < RR; ++r) {
                    du[r].x += v[u][r].x, du[r].y += v[u][r].y;
                    s.x += v[u][r].x, s.y += v[u][r].y;
                }
                if (l + u < l1) dv[(int64_t)(l + u) * (C / 2)] = s;
            }
        }
    } else {
        for (int l = l0; l < l1; ++l) {
            double2 v[GS_ROWS];
#pragma unroll
            for (int r = 0; r < GS_ROWS; ++r)
                v[r] = r < nr ? __ldcs(base + ((int64_t)r * L + l) * (C / 2)) : make_double2(0.0, 0.0);
            double2 s = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < GS_ROWS; ++r) {
                du[r].x += v[r].x, du[r].y += v[r].y;
                s.x += v[r].x, s.y += v[r].y;
            }
            dv[(int64_t)l * (C / 2)] = s;
        }
    }
    double2* duo = reinterpret_cast<double2*>(dUpart) + (int64_t)blockIdx.y * n * (C / 2);
#pragma unroll
    for (int r = 0; r < GS_ROWS; ++r)
        if (r < nr) duo[(int64_t)(r0 + r) * (C / 2) + c2] = du[r];
}

static int gs_sm_count()
{
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    return n_sm;
}

// Number of row blocks (= leading dimension of dVpart) drp_gru_input_sums_f64 uses for n rows.
DRP_API int drp_gru_input_sums_blocks(int n)
{
    if (n <= 0) return 0;
    const int n_sm = gs_sm_count();
    int rows = (n + n_sm - 1) / n_sm;
    if (rows > GS_ROWS) rows = GS_ROWS;
    return (n + rows - 1) / rows;
}

// Number of position chunks (= leading dimension of dUpart): enough CTAs for two per SM, at least 32 positions each.
DRP_API int drp_gru_input_sums_chunks(int n, int L)
{
    if (n <= 0 || L <= 0) return 0;
    const int nblk = drp_gru_input_sums_blocks(n);
    int want = (2 * gs_sm_count() + nblk - 1) / nblk;
    const int most = (L + 31) / 32;
    if (want > most) want = most;
    if (want > 16) want = 16;
    return want < 1 ? 1 : want;
}

// dUpart [drp_gru_input_sums_chunks(n, L), n, C] (the caller adds the chunks), dVpart [drp_gru_input_sums_blocks(n), L, C].
DRP_API int drp_gru_input_sums_f64(const double* dgi, int n, int L, int C, double* dUpart, double* dVpart, void* stream)
{
    if (!dgi) return -1;
    if (n <= 0 || L <= 0) return -2;
    if (C <= 0 || (C & 1) || C / 2 > 256) return -4;
    if (!dUpart || !dVpart) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = drp_gru_input_sums_blocks(n);
    const int rows = (n + nblk - 1) / nblk;
    const int nch = drp_gru_input_sums_chunks(n, L);
    int lper = (L + nch - 1) / nch;
    lper = (lper + GS_LU - 1) / GS_LU * GS_LU;
    DRP_LAUNCH(DRP_K_GRU_SUMS, (double)n * L * C * 8.0 + (double)nblk * L * C * 8.0, st);
    gru_input_sums_kernel<<<dim3(nblk, nch), 256, 0, st>>>(dgi, n, L, C, rows, lper, dUpart, dVpart);
    return drp_launch_status();
}

// Workspace (doubles) of drp_gru_wgrad_f64.
DRP_API int64_t drp_gru_wgrad_workspace_elems(void)
{
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    return (int64_t)n_sm * 2 * (GW_MT * 8) * 64;
}

// dwhh [2,3H,H], dbhh [2,3H] from dgi [n,L,2,3H] (drp_gru_bwd_f64), the saved gates, the hidden states and h0.
static int gru_wgrad_dispatch(const double* dgi, const double* gates, const double* out, const double* h0, int n, int L,
                              int H, double* dwhh, double* dbhh, double* ws, int64_t ws_elems, int ndir, void* stream)
{
    if (!dgi || !gates || !out || !h0) return -1;
    if (n <= 0 || L <= 0) return -5;
    if (H != GH) return -7;
    if (!dwhh || !dbhh) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    static int n_sm = 148;
    if (drp_first_on_device(30)) {
        cudaFuncSetAttribute(gru_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GW_SMEM);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t nblk = ((int64_t)n * L + GW_PAIRS - 1) / GW_PAIRS;
    int nchunks = n_sm / ndir > 0 ? n_sm / ndir : 1;
    if (nchunks > nblk) nchunks = (int)nblk;
    if (!ws || ws_elems < (int64_t)nchunks * 2 * (GW_MT * 8) * 64) return -10;
    {
        DRP_LAUNCH(DRP_K_GRU_WGRAD, 2.0 * (double)n * L * ndir * G3 * (GH + 1), st);
        gru_wgrad_kernel<<<ndir * nchunks, GW_THREADS, GW_SMEM, st>>>(dgi, gates, out, h0, n, L, ws, ndir, ndir);
    }
    {
        DRP_LAUNCH(DRP_K_GRU_WGRAD, (double)nchunks * ndir * G3 * (GH + 1) * 8.0, st);
        gru_wgrad_reduce_kernel<<<(ndir * G3 * (GH + 1) + 255) / 256, 256, 0, st>>>(ws, nchunks, dwhh, dbhh, ndir);
    }
    return drp_launch_status();
}

DRP_API int drp_gru_wgrad_f64(const double* dgi, const double* gates, const double* out, const double* h0, int n, int L,
                              int H, double* dwhh, double* dbhh, double* ws, int64_t ws_elems, void* stream)
{
    return gru_wgrad_dispatch(dgi, gates, out, h0, n, L, H, dwhh, dbhh, ws, ws_elems, 2, stream);
}

// Direction 0 only: dgi0 [n,L,3H] compact -> dwhh[0], dbhh[0] (the direction-1 slices of the outputs are not written).
DRP_API int drp_gru_wgrad_dir0_f64(const double* dgi0, const double* gates, const double* out, const double* h0, int n, int L,
                                   int H, double* dwhh, double* dbhh, double* ws, int64_t ws_elems, void* stream)
{
    return gru_wgrad_dispatch(dgi0, gates, out, h0, n, L, H, dwhh, dbhh, ws, ws_elems, 1, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Tall-skinny Gram product  C [M,N] = A^T B,  s [M] = column sums of A,  for A [K,M], B [K,N] with K in the millions and
// M, N <= 192: the weight / bias gradients of the dense layers around the recurrences (autograd of nn.Linear:
// dW = dY^T X, db = sum dY; S/models_utils.py:54-65,93-99).  cuBLAS runs these "tn" shapes without split-K at a fraction
// of the HBM rate.  Here: split-K over one CTA per SM, 32-row blocks of both operands arrive as ONE TMA bulk copy each
// (rows are contiguous), fp64 mma.sync accumulators in registers, deterministic second-stage sum.
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int GG_STAGES = 3;
constexpr int GG_MI = 3, GG_NI = 5;              // accumulator tiles (8x8) per warp: a block of up to 3 x 5

// How the mt_n x nt_n grid of 8x8 output tiles is cut into 8 rectangular blocks, one per warp: wm_n warps along m (8, 4, 2
// or 1), 8 / wm_n along n, each block at most GG_MI x GG_NI tiles; 0 if no cut fits.  Among the cuts that fit, the one with
// the fewest fragment loads per k-step (mi + ni).
__host__ __device__ inline int gg_warp_cut(int mt_n, int nt_n)
{
    int best = 0, cost = 1 << 30;
    for (int wm = 8; wm >= 1; wm >>= 1) {
        const int wn = 8 / wm, mi = (mt_n + wm - 1) / wm, ni = (nt_n + wn - 1) / wn;
        if (mi <= GG_MI && ni <= GG_NI && mi + ni < cost) best = wm, cost = mi + ni;
    }
    return best;
}
// K-rows per pipeline stage: ~48 KB per stage (three stages in flight), at least 32 and at most 128 rows, a multiple of 16.
// With narrow operands (21 + 21 columns) 32-row stages are 10 KB and the loop is bound by its per-iteration overhead.
__host__ __device__ inline int gg_stage_rows(int M, int N)
{
    int r = (48 * 1024) / ((M + N) * 8);
    r = r > 128 ? 128 : (r < 32 ? 32 : r);
    return r & ~15;
}

// Every warp owns a RECTANGULAR block of tiles, so that per k-step it loads mi A-fragments and ni B-fragments for mi x ni
// tensor instructions (the first version dealt tiles round-robin: two shared-memory loads per instruction, and shared-memory
// bandwidth - 283 KB per 32-row block at M = 180 - was a third bound next to the DMMA pipe and HBM, all three ~2.2 k cycles
// per block, 4.9 k measured).
template <int MI, int NI>
__global__ void __launch_bounds__(GTHREADS, 1) gram_tn_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                              int64_t K, int M, int N, double* __restrict__ part, int rows,
                                                              int wm_n)
{
    extern __shared__ __align__(16) uint8_t gsm[];
    const int NE = N + 1;                                        // + the ones column (column sums of A)
    const int mt_n = (M + 7) / 8, nt_n = (NE + 7) / 8;
    const int stage_elems = rows * (M + N);
    double* ring = reinterpret_cast<double*>(gsm);
    const uint32_t bars = g_smem_u32(gsm + (size_t)GG_STAGES * stage_elems * 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int wn_n = 8 / wm_n;
    const int mt0 = (warp / wn_n) * MI, nt0 = (warp % wn_n) * NI;   // this warp's MI x NI block of tiles
    bool vm[MI], vn[NI];                                         // warp-uniform: which tiles of the block exist (edge warps)
#pragma unroll
    for (int i = 0; i < MI; ++i) vm[i] = mt0 + i < mt_n;
#pragma unroll
    for (int j = 0; j < NI; ++j) vn[j] = nt0 + j < nt_n;
    const int64_t nblk = (K + rows - 1) / rows;
    const int64_t b0 = nblk * blockIdx.x / gridDim.x, b1 = nblk * (blockIdx.x + 1) / gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < GG_STAGES; ++s) g_mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t blk) {                              // one thread: two contiguous bulk copies
        const int s = (int)((blk - b0) % GG_STAGES);
        if (K - blk * rows < rows) return;                       // ragged last block: loaded with plain loads by the consumer
        double* dst = ring + (int64_t)s * stage_elems;
        g_mbar_expect_tx(bars + 8 * s, (uint32_t)rows * (M + N) * 8);
        g_bulk_g2s(g_smem_u32(dst), A + blk * rows * M, (uint32_t)rows * M * 8, bars + 8 * s);
        g_bulk_g2s(g_smem_u32(dst + rows * M), B + blk * rows * N, (uint32_t)rows * N * 8, bars + 8 * s);
    };
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (tid == 0)
        for (int64_t blk = b0; blk < min(b1, b0 + GG_STAGES - 1); ++blk) issue(blk);
    for (int64_t blk = b0; blk < b1; ++blk) {
        const int it = (int)(blk - b0);
        const int s = it % GG_STAGES;
        if (tid == 0 && blk + GG_STAGES - 1 < b1) issue(blk + GG_STAGES - 1);    // its stage was drained in iteration it-1
        double* sa = ring + (int64_t)s * stage_elems;
        double* sb = sa + rows * M;
        const int valid = (int)min((int64_t)rows, K - blk * rows);
        if (valid == rows) {
            g_mbar_wait(bars + 8 * s, (uint32_t)(it / GG_STAGES) & 1u);
        } else {                                                 // ragged last block (no 16-byte size guarantee): plain loads, zero fill
            const double* ga = A + blk * rows * M;
            const double* gb = B + blk * rows * N;
            for (int idx = tid; idx < rows * M; idx += GTHREADS) sa[idx] = idx < valid * M ? ga[idx] : 0.0;
            for (int idx = tid; idx < rows * N; idx += GTHREADS) sb[idx] = idx < valid * N ? gb[idx] : 0.0;
            __syncthreads();
        }
#pragma unroll 4
        for (int ks = 0; ks < rows / 4; ++ks) {                  // rows is a multiple of 16
            const int k = 4 * ks + tig;
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int m = 8 * (mt0 + i) + gid;
                a[i] = (vm[i] && m < M) ? sa[k * M + m] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int nn = 8 * (nt0 + j) + gid;
                b[j] = !vn[j] ? 0.0 : (nn < N ? sb[k * N + nn] : (nn == N && k < valid ? 1.0 : 0.0));
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);   // tiles that do not exist add zeros
        }
        __syncthreads();                                          // stage s drained before it is refilled
    }
    double* dst = part + (int64_t)blockIdx.x * (mt_n * 8) * (nt_n * 8);
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
            if (vm[i] && vn[j])
                *reinterpret_cast<double2*>(dst + (int64_t)(8 * (mt0 + i) + gid) * (nt_n * 8) + 8 * (nt0 + j) + 2 * tig) =
                    make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---------------------------------------------------------------------------------------------------------------
// Tall-skinny dense layer  Y [M, N] = X [M, K] W^T + bias  for M in the millions and K, N <= 192: the input projection of
// the encoder GRU (K 21 -> N 180), the collapsed fc2.fc1 maps, the decoder head (120 -> 21) and their input gradients
// (dY W).  cuBLAS runs these fp64 shapes at 10-25 % of what the memory traffic allows (cutlass_80 d884 tiles sized for
// square problems).  Here W^T stays in shared memory and every WARP is its own pipeline: it owns a contiguous range of
// 8-row tiles of X, streams them through a private TMA-bulk ring (the 8 rows of a tile are contiguous: one copy, issued
// by the warp's lane 0), turns a tile into all N outputs with fp64 mma.sync (accumulators start from the bias) and
// stores them; no CTA-wide barrier in the loop, so one warp's stores overlap the others' arithmetic.  HBM-bound by
// construction: X read once, Y written once.
//   Wt [K, N] row-major (= W^T), bias [N] nullable.  NT = number of 8-column output tiles (3, 15 or 23).
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int tg_ns(int nt) { return nt * 8 + 4; }             // W^T row stride: == 12 (mod 16), conflict-free
__host__ __device__ constexpr int tg_wbytes(int K, int nt) { return ((K + 3) / 4 * 4) * tg_ns(nt) * 8; }
__host__ __device__ constexpr int tg_obytes(int nt, int mtw) { return 8 * 8 * mtw * (nt * 8) * 8; }   // 8 warps x one output tile
__host__ __device__ constexpr int tg_stages(int K, int nt, int mtw)   // ring stages per warp that fit next to W^T and the output tiles
{
    const int room = (227 * 1024 - 1024 - tg_wbytes(K, nt) - tg_obytes(nt, mtw)) / (8 * 8 * mtw * K * 8);
    return room > 4 ? 4 : room;
}
__host__ __device__ constexpr int tg_smem(int K, int nt, int mtw)
{
    return tg_wbytes(K, nt) + tg_obytes(nt, mtw) + tg_stages(K, nt, mtw) * 8 * 8 * mtw * K * 8 + 1024;
}

// MTW = 8-row tiles per warp iteration: the B fragments (W^T from shared memory, one 8-byte load per DMMA) are shared by
// the MTW row tiles, which halves the shared-memory traffic that otherwise rivals the DMMA time (138 loads per 138 DMMAs).
template <int NT, int MTW>
__global__ void __launch_bounds__(GTHREADS, 1) tall_gemm_kernel(const double* __restrict__ X, const double* __restrict__ Wt,
                                                                const double* __restrict__ bias, int64_t M, int K, int N,
                                                                double* __restrict__ Y)
{
    constexpr int NS = tg_ns(NT), TR = 8 * MTW;                                       // rows per warp iteration
    extern __shared__ __align__(16) uint8_t gsm[];
    const int Kp = (K + 3) / 4 * 4, stages = tg_stages(K, NT, MTW);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    double* wsm = reinterpret_cast<double*>(gsm);                                    // [Kp][NS]
    double* osm = wsm + Kp * NS + warp * TR * (NT * 8);                               // this warp's output tile [TR][N] (dense)
    double* ring = wsm + Kp * NS + 8 * TR * (NT * 8) + (int64_t)warp * stages * TR * K;   // this warp's [stages][TR][K]
    const uint32_t bars = g_smem_u32(gsm + tg_smem(K, NT, MTW) - 1024) + warp * 4 * 8;  // this warp's barriers
    for (int idx = tid; idx < Kp * NS; idx += GTHREADS) {
        const int k = idx / NS, c = idx - k * NS;
        wsm[idx] = (k < K && c < N) ? Wt[(int64_t)k * N + c] : 0.0;
    }
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) g_mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();                                             // W^T in place (the only CTA-wide barrier)
    const int64_t ntiles = (M + TR - 1) / TR;
    const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    const int64_t t0 = ntiles * gw / nw, t1 = ntiles * (gw + 1) / nw;
    const uint32_t tile_bytes = (uint32_t)TR * K * 8, out_bytes = (uint32_t)TR * N * 8;
    auto issue = [&](int64_t tile) {                             // lane 0: the rows of a tile are one contiguous copy
        if (M - tile * TR < TR) return;                          // ragged last tile: plain loads below
        const int s = (int)((tile - t0) % stages);
        g_mbar_expect_tx(bars + 8 * s, tile_bytes);
        g_bulk_g2s(g_smem_u32(ring + (int64_t)s * TR * K), X + tile * TR * K, tile_bytes, bars + 8 * s);
    };
    if (lane == 0)
        for (int64_t tile = t0; tile < min(t1, t0 + stages - 1); ++tile) issue(tile);
    const bool vec = (N & 1) == 0;                               // 16-byte shared-memory stores need an even row length
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int it = (int)(tile - t0);
        const int s = it % stages;
        if (tile + stages - 1 < t1) {                            // refill the stage this warp drained in the last iteration
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(tile + stages - 1);
            }
        }
        double* xs = ring + (int64_t)s * TR * K;
        const int valid = (int)min((int64_t)TR, M - tile * TR);
        if (valid == TR) {
            g_mbar_wait(bars + 8 * s, (uint32_t)(it / stages) & 1u);
        } else {
            const double* gx = X + tile * TR * K;
            for (int idx = lane; idx < TR * K; idx += 32) xs[idx] = idx < valid * K ? gx[idx] : 0.0;
            __syncwarp();
        }
        double acc[MTW][NT][2];
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int ks = 0; ks < Kp / 4; ++ks) {
            const int k = 4 * ks + tig;
            double a[MTW];
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) a[mt] = k < K ? xs[(8 * mt + gid) * K + k] : 0.0;
            const double* wr = wsm + k * NS + gid;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b = wr[8 * nt];
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b);
            }
        }
        // the output rows of a tile are contiguous in Y: staged densely in shared memory and written by ONE bulk store (the
        // accumulator fragments would otherwise leave as 64-byte pieces of 8 different rows per instruction)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile has left osm
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = 8 * nt + 2 * tig;
            const double b0v = (bias && c < N) ? __ldg(bias + c) : 0.0, b1v = (bias && c + 1 < N) ? __ldg(bias + c + 1) : 0.0;
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                double* orow = osm + (8 * mt + gid) * N;
                if (vec) {
                    if (c < N) *reinterpret_cast<double2*>(orow + c) = make_double2(acc[mt][nt][0] + b0v, acc[mt][nt][1] + b1v);
                } else {
                    if (c < N) orow[c] = acc[mt][nt][0] + b0v;
                    if (c + 1 < N) orow[c + 1] = acc[mt][nt][1] + b1v;
                }
            }
        }
        __syncwarp();
        if (valid == TR) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Y + tile * TR * N),
                             "r"(g_smem_u32(osm)), "r"(out_bytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int idx = lane; idx < valid * N; idx += 32) Y[tile * TR * N + idx] = osm[idx];
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // stores complete before the CTA exits
}

__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ part, int nparts, int M, int N,
                                                          double* __restrict__ C, double* __restrict__ colsum)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // over [M][N+1]
    if (idx >= M * (N + 1)) return;
    const int m = idx / (N + 1), j = idx - m * (N + 1);
    const int ldp = ((N + 1 + 7) / 8) * 8, rows = ((M + 7) / 8) * 8;
    double s = 0.0;
    for (int k = 0; k < nparts; ++k) s += part[((int64_t)k * rows + m) * ldp + j];
    if (j < N) C[(int64_t)m * N + j] = s;
    else if (colsum) colsum[m] = s;
}

}  // namespace

DRP_API int64_t drp_gram_workspace_elems(int M, int N)
{
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    return (int64_t)n_sm * (((M + 7) / 8) * 8) * (((N + 1 + 7) / 8) * 8);
}

// Does drp_gram_tn_f64 take the shape (M, N)?  (The 8x8 output tiles must split into 8 blocks of at most 3 x 5 and three
// pipeline stages must fit in shared memory.)
DRP_API int drp_gram_tn_supported(int M, int N)
{
    if (M <= 0 || N <= 0 || M > 192 || N > 192) return 0;
    if (gg_warp_cut((M + 7) / 8, (N + 1 + 7) / 8) == 0) return 0;
    return GG_STAGES * gg_stage_rows(M, N) * (M + N) * 8 + 64 <= 200 * 1024;
}

// C [M,N] = A^T B and colsum [M] = sum_k A[k][:] (nullable) for A [K,M], B [K,N], both row-major contiguous fp64.
DRP_API int drp_gram_tn_f64(const double* A, const double* B, int64_t K, int M, int N, double* C, double* colsum, double* ws,
                            int64_t ws_elems, void* stream)
{
    if (!A || !B) return -1;
    if (K <= 0) return -3;
    if (M <= 0 || N <= 0 || M > 192 || N > 192) return -4;
    if (!C) return -6;
    const int wm_n = gg_warp_cut((M + 7) / 8, (N + 1 + 7) / 8);
    if (wm_n == 0) return -4;
    const int rows = gg_stage_rows(M, N);
    cudaStream_t st = (cudaStream_t)stream;
    static int n_sm = 148;
    if (drp_first_on_device(57)) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const int wn_n = 8 / wm_n, mi = ((M + 7) / 8 + wm_n - 1) / wm_n, ni = ((N + 1 + 7) / 8 + wn_n - 1) / wn_n;
    void (*kern)(const double*, const double*, int64_t, int, int, double*, int, int) = nullptr;
#define GG_CASE(a, b) if (mi == a && ni == b) kern = gram_tn_kernel<a, b>;
    GG_CASE(1, 1) GG_CASE(1, 2) GG_CASE(1, 3) GG_CASE(1, 4) GG_CASE(1, 5)
    GG_CASE(2, 1) GG_CASE(2, 2) GG_CASE(2, 3) GG_CASE(2, 4) GG_CASE(2, 5)
    GG_CASE(3, 1) GG_CASE(3, 2) GG_CASE(3, 3) GG_CASE(3, 4) GG_CASE(3, 5)
#undef GG_CASE
    if (!kern) return -4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // idempotent, ~1 us
    const int smem = GG_STAGES * rows * (M + N) * 8 + 64;
    if (smem > 200 * 1024) return -4;
    const int64_t nblk = (K + rows - 1) / rows;
    const int grid = (int)(nblk < n_sm ? nblk : n_sm);
    if (!ws || ws_elems < (int64_t)grid * (((M + 7) / 8) * 8) * (((N + 1 + 7) / 8) * 8)) return -8;
    {
        DRP_LAUNCH(DRP_K_GRAM, (double)K * (M + N) * 8.0, st);
        kern<<<grid, GTHREADS, smem, st>>>(A, B, K, M, N, ws, rows, wm_n);
    }
    {
        DRP_LAUNCH(DRP_K_GRAM, (double)grid * M * (N + 1) * 8.0, st);
        gram_reduce_kernel<<<(M * (N + 1) + 255) / 256, 256, 0, st>>>(ws, grid, M, N, C, colsum);
    }
    return drp_launch_status();
}

template <int NT, int MTW>
static void tall_gemm_launch(const double* X, const double* Wt, const double* bias, int64_t M, int K, int N, double* Y,
                             int n_sm, cudaStream_t st)
{
    const int smem = tg_smem(K, NT, MTW);
    cudaFuncSetAttribute(tall_gemm_kernel<NT, MTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int64_t ntiles = (M + 8 * MTW - 1) / (8 * MTW);
    const int64_t want = (ntiles + 7) / 8;                        // at least one tile per warp
    const int grid = (int)(want < n_sm ? want : n_sm);
    DRP_LAUNCH(DRP_K_TALL_GEMM, (double)M * (K + N) * 8.0, st);
    tall_gemm_kernel<NT, MTW><<<grid, GTHREADS, smem, st>>>(X, Wt, bias, M, K, N, Y);
}

// Y [M,N] = X [M,K] Wt + bias for row-major contiguous fp64 X, Wt [K,N] (the transposed weight), bias [N] (nullable).
// K <= 192, N <= 184.  Returns 1 (nothing launched) when W^T and the ring do not fit in shared memory (large K and N).
DRP_API int drp_tall_gemm_f64(const double* X, const double* Wt, const double* bias, int64_t M, int K, int N, double* Y,
                              void* stream)
{
    if (!X || !Wt) return -1;
    if (M <= 0) return -4;
    if (K <= 0 || K > 192) return -5;
    if (N <= 0 || N > 184) return -6;
    if (!Y) return -7;
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0 || (reinterpret_cast<uintptr_t>(Y) & 15) != 0) return -1;
    const int nt = N <= 24 ? 3 : (N <= 120 ? 15 : 23);
    const int mtw = tg_stages(K, nt, 2) >= 2 ? 2 : 1;             // two row tiles per warp share the B fragments when shared memory allows
    if (tg_stages(K, nt, mtw) < 1) return 1;                      // declined: W^T + ring do not fit
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mtw == 2) {
        if (nt == 3) tall_gemm_launch<3, 2>(X, Wt, bias, M, K, N, Y, n_sm, st);
        else if (nt == 15) tall_gemm_launch<15, 2>(X, Wt, bias, M, K, N, Y, n_sm, st);
        else tall_gemm_launch<23, 2>(X, Wt, bias, M, K, N, Y, n_sm, st);
    } else {
        if (nt == 3) tall_gemm_launch<3, 1>(X, Wt, bias, M, K, N, Y, n_sm, st);
        else if (nt == 15) tall_gemm_launch<15, 1>(X, Wt, bias, M, K, N, Y, n_sm, st);
        else tall_gemm_launch<23, 1>(X, Wt, bias, M, K, N, Y, n_sm, st);
    }
    return drp_launch_status();
}
