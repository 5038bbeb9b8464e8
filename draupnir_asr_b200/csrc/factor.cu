// Batched blocked partial Cholesky — the engine behind the OU prior, its gradients and the conditional posterior.
//
// Replaces (reference, S/ = /root/reference/draupnir/src/draupnir/):
//   torch.linalg.cholesky inside MultivariateNormal.__init__      S/models.py:118
//   solve_triangular in MultivariateNormal.log_prob               S/models.py:118 (torch _batch_mahalanobis)
//   autograd cholesky_backward / triangular_solve_backward        (K5 in SURVEY.md §2b)
//   torch.linalg.inv x3 + matmul x2 in conditional_sampling*      S/models.py:167,188,190,193,214,235,237,239
//
// One primitive does all of it: eliminate the first `nfactor` columns of a lower-stored symmetric matrix
//
//            [ K    .    . ]   rows 0..n-1     (covariance)
//        M = [ x^T  0    . ]   row  n          (right-hand side)
//            [ B    0    C ]   rows n+1..      (B = I for the prior, K_ab for the conditional; C = 0 or K_aa)
//
// which leaves  row n        = y^T = (L^-1 x)^T,        M[n][n]      = -y^T y              (Mahalanobis)
//               M[n+1+i][n]  = -(B K^-1 x)_i            (= -alpha_i or -mu_i)
//               M[n+1+i][n+1+j] = C - B K^-1 B^T        (= -K^-1   or  Sigma_a|b)
// i.e. the solve, the inverse and the Schur complement are all by-products of the same trailing updates, and the
// only kernels are potf2 (diagonal block), trsm (panel rows) and syrk (trailing update).  With B = I the rows of
// B below the current panel are still exactly zero, so the active extent grows with the panel (`grow_base`) and
// the flop count is the optimal n^3/3 + 2n^3/3.
//
// Layout: row-major, leading dimension `ld`, batch stride `bs` (elements); only entries with col <= row are
// read or written.  fp64 throughout (the reference dtype).  The trailing update runs on the fp64 tensor pipe
// (mma.sync m8n8k4); see DESIGN.md for the tcgen05 split-integer variant.
#include "common.cuh"
#include "factor.cuh"
#include "instrument.cuh"
#include "slice.cuh"

namespace {

constexpr int NB  = DRP_NB;    // panel width
constexpr int PLD = NB + 1;    // padded row stride of smem panels (conflict-free thread-per-row access)
constexpr int PR  = 128;       // rows per CTA in trsm
constexpr int TS  = 64;        // syrk tile
constexpr int SLD = 68;        // syrk smem row stride: (row*68 + tig) mod 16 distinct over a half warp
static_assert(PR == 2 * NB, "trsm staging assumes two rows per pass");

// ---------------------------------------------------------------------------------------------------------------
// potf2: factor the nb x nb diagonal block at (c0, c0).  One CTA of 256 threads per matrix; thread (ty, tx) keeps the
// 4x4 block (rows 4ty.., columns 4tx..) in registers, tid = 16 tx + ty so that the 16 owners of a 4-column panel are
// one half-warp.  Per panel: the owners factor the 64x4 panel among themselves with warp shuffles (pivot and the three
// panel-row broadcasts come from the lane that owns the diagonal block), publish it to shared memory, and after ONE
// barrier every thread applies the rank-4 update to its block.  Entries right of the diagonal collect garbage that is
// never read back nor stored.  (A fully unrolled thread-per-row variant was instruction-fetch bound, a column-at-a-time
// variant paid two barriers per column.)
// ---------------------------------------------------------------------------------------------------------------
// 1 / sqrt(x) for x > 0: MUFU.RSQ64H seed + two Newton steps, branch-free (the library rsqrt carries a slow-path CALL in the
// middle of the 64-step dependent chain of the panel factorisation).  Non-positive / non-finite pivots produce NaN or inf as
// the library routine does; the caller flags them through `info` before the value is used.
__device__ __forceinline__ double potf2_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

// 1 / x for x > 0: MUFU.RCP64H seed (>= 20 bits) and ONE fourth-order correction, y (1 + e)(1 + e^2) with e = 1 - x y: three
// dependent fp64 operations behind the seed.  This is the chain every column of the panel factorisation waits for.
__device__ __forceinline__ double potf2_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double y1 = fma(y, e, y), e2 = e * e;
    return fma(y1, e2, y1);
}

// Factor the register-distributed 64 x 64 block `a` (thread (ty, tx) = (tid & 15, tid >> 4) holds rows 4 ty.., columns 4 tx..)
// in place.  `*sbad` (shared, zeroed by the caller before a barrier) receives 1 + the first column with a non-positive pivot.
// Ends with a barrier.
// The panel in shared memory: row r of the 64 x 4 panel at LP_ROW(r) doubles - 32 bytes per row plus 16 bytes after every
// fourth row.  The rank-4 update reads rows 4 ty .. 4 ty + 3 with 16-byte loads, lane = ty: at a plain 32-byte row stride the
// eight lanes of a quarter-warp are 128 bytes apart, i.e. on the SAME four banks (a 8-way conflict on each of the 8 loads
// per thread and panel: the ncu source page showed the update phase waiting on shared memory); with the shift they are 144
// bytes apart and cover all 32 banks.
#define LP_ROW(r) ((r) * 4 + 2 * ((r) >> 2))
constexpr int LP_ELEMS = NB * 4 + 2 * (NB / 4);

__device__ __forceinline__ void potf2_regs(double (&a)[4][4], double* Lp, int* sbad)
{
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15, lane = tid & 31;
    for (int cb = 0; cb < NB / 4; ++cb) {
        if (tx == cb) {                                            // one half-warp: lanes (cb & 1) * 16 + ty
            const unsigned mask = (cb & 1) ? 0xffff0000u : 0x0000ffffu;
            const int lane_d = (lane & 16) + cb;                   // owner of the diagonal block (ty == cb)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = 4 * cb + cc;
                const double piv = __shfl_sync(mask, a[cc][cc], lane_d);
                double ack[4];                                     // unscaled column entries A[4 cb + k][c]: on their way
#pragma unroll                                                     // before either chain below has finished
                for (int k = cc + 1; k < 4; ++k) ack[k] = __shfl_sync(mask, a[k][cc], lane_d);
                if (ty == 0 && !(piv > 0.0) && *sbad == 0) *sbad = c + 1;
                // two independent chains from the pivot: 1 / piv, which the remaining columns of the panel wait for
                // (a_rk -= (a_rc / piv) a_kc: seed + 3 + 2 dependent operations per column), and 1 / sqrt(piv) for the stored
                // column of L, which nothing in the panel waits for
                const double ip = potf2_rcp(piv);
                const double rd = potf2_rsqrt(piv);
#pragma unroll
                for (int k = cc + 1; k < 4; ++k) {                 // remaining columns of the panel (rows <= c: never read)
#pragma unroll
                    for (int r = 0; r < 4; ++r) a[r][k] = fma(-(a[r][cc] * ip), ack[k], a[r][k]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = 4 * ty + r;
                    const double lv = i > c ? a[r][cc] * rd : (i == c ? piv * rd : 0.0);
                    if (i >= c) a[r][cc] = lv;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) Lp[LP_ROW(4 * ty + r) + k] = (4 * ty + r >= 4 * cb + k) ? a[r][k] : 0.0;
        }
        __syncthreads();
        if (tx > cb) {                                             // rank-4 update of the trailing blocks
            double lr[4][4], lc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    lr[r][q] = Lp[LP_ROW(4 * ty + r) + q];
                    lc[r][q] = Lp[LP_ROW(4 * tx + r) + q];
                }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double v = a[r][k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v = fma(-lr[r][q], lc[k][q], v);
                    a[r][k] = v;
                }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) potf2_kernel(double* __restrict__ Mb, int64_t ld, int64_t bs, int c0, int nb,
                                                    int32_t* __restrict__ info)
{
    __shared__ __align__(16) double Lp[LP_ELEMS];   // the current panel of L
    __shared__ int sbad;
    double* A = Mb + (int64_t)blockIdx.x * bs + (int64_t)c0 * ld + c0;
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15;
    if (tid == 0) sbad = 0;
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * ty + r, j = 4 * tx + k;
            // rows / columns beyond nb: identity, so the elimination is a no-op there
            a[r][k] = (i < nb && j <= i) ? A[(int64_t)i * ld + j] : (i == j ? 1.0 : 0.0);
        }
    __syncthreads();
    potf2_regs(a, Lp, &sbad);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * ty + r, j = 4 * tx + k;
            if (i < nb && j <= i) A[(int64_t)i * ld + j] = a[r][k];
        }
    if (tid == 0 && sbad && sbad <= nb && info) atomicCAS(info + blockIdx.x, 0, c0 + sbad);
}

// ---------------------------------------------------------------------------------------------------------------
// trsm: rows r in [c0+nb, R): X[r][c0:c0+nb] <- X L_jj^-T.  One thread per row with the row in REGISTERS (fully
// unrolled forward substitution, two independent accumulation chains); the factored block is read from shared memory
// with warp-uniform (broadcast) 16-byte loads; rows are staged through shared memory for coalesced global access.
// ---------------------------------------------------------------------------------------------------------------
// Optional fused slicing (`sl.sl != nullptr`; only for nb == NB): the solved 64 columns of rows [c0 + 64, R) ARE the panel of
// the in-panel K = 64 trailing update that follows, so its int8 slices (ozaki.cu: same fixed-point split, same 128-byte
// swizzled K-major image, bit-identical to oz_slice_kernel) are written from shared memory while the rows are on chip,
// instead of by a separate kernel that re-reads them (16 launches of ~40 us per 2,000-leaf factorisation, latency-bound).
struct TrsmSlice {
    uint8_t* sl;        // nullptr: plain solve
    int64_t sl_bs;
    int32_t* scl;
    int* counter;
    int Rpad, S;
    int KBN;            // 128-column k-blocks of the sliced panel (panel_solve_kernel only; trsm_kernel slices 64 columns)
};

__global__ void __launch_bounds__(PR) trsm_kernel(double* __restrict__ Mb, int64_t ld, int64_t bs, int c0, int nb, int R,
                                                  const TrsmSlice sl)
{
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                  // [NB][NB] factored diagonal block (identity beyond nb)
    double* invd = S + NB * NB;      // [NB]
    double* X = invd + NB;           // [PR][PLD]
    double* A = Mb + (int64_t)blockIdx.y * bs;
    const int tid = threadIdx.x;
    const int r0 = c0 + nb + blockIdx.x * PR;
    const int nrows = min(PR, R - r0);
    {   // two rows per pass (PR = 2 NB threads), 16 passes in flight at a time
        const int j = tid & (NB - 1), half = tid >> 6;
        for (int i0 = 0; i0 < NB; i0 += 32) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int i = i0 + 2 * u + half;
                v[u] = (i < nb && j <= i) ? A[(int64_t)(c0 + i) * ld + c0 + j] : (i == j ? 1.0 : 0.0);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) S[(i0 + 2 * u + half) * NB + j] = v[u];
        }
        for (int i0 = 0; i0 < PR; i0 += 32) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int i = i0 + 2 * u + half;
                v[u] = (i < nrows && j < nb) ? A[(int64_t)(r0 + i) * ld + c0 + j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) X[(i0 + 2 * u + half) * PLD + j] = v[u];
        }
    }
    __syncthreads();
    if (tid < NB) invd[tid] = 1.0 / S[tid * NB + tid];
    __syncthreads();
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = X[tid * PLD + j];
    // x_c = (x_c - sum_{p<c} x_p L[c][p]) / L[c][c]: the sum runs in FOUR independent accumulation chains (the solve is a
    // pure dependent-FMA latency chain per thread at 8 resident warps per SM; two chains took 1.35 ms per 2,000-leaf step)
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const double* lc = S + c * NB;
        double a0 = x[c], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int p = 0; p + 3 < c; p += 4) {
            const double2 l01 = *reinterpret_cast<const double2*>(lc + p);
            const double2 l23 = *reinterpret_cast<const double2*>(lc + p + 2);
            a0 = fma(-x[p], l01.x, a0);
            a1 = fma(-x[p + 1], l01.y, a1);
            a2 = fma(-x[p + 2], l23.x, a2);
            a3 = fma(-x[p + 3], l23.y, a3);
        }
#pragma unroll
        for (int p = c & ~3; p < c; ++p) a0 = fma(-x[p], lc[p], a0);
        x[c] = ((a0 + a1) + (a2 + a3)) * invd[c];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) X[tid * PLD + j] = x[j];
    __syncthreads();
    {
        const int j = tid & (NB - 1), half = tid >> 6;
        if (j < nb) {
#pragma unroll 8
            for (int i = half; i < PR; i += 2)
                if (i < nrows) A[(int64_t)(r0 + i) * ld + c0 + j] = X[i * PLD + j];
        }
    }
    if (sl.sl == nullptr) return;
    // ---- slices of this CTA's 128 rows (one 128-row strip of the update; rows >= R are zeros): two rows per warp iteration,
    //      16 lanes x 4 consecutive columns each, 4-byte stores (64 contiguous bytes per row and slice)
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) sl.counter[0] = sl.counter[1] = 0;
    const int warp = tid >> 5, lane = tid & 31, hl = lane & 15, sub = lane >> 4;
    const int NS = sl.S;
    uint8_t* dst = sl.sl + (int64_t)blockIdx.y * sl.sl_bs;
    for (int it = 0; it < PR / 8; ++it) {
        const int i = warp * (PR / 4) + 2 * it + sub;            // row inside the CTA
        const int rr = blockIdx.x * PR + i;                      // row of the slice buffer (relative to c0 + nb)
        double xv[4];
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xv[j] = X[i * PLD + 4 * hl + j];                     // rows >= nrows were staged as zeros and stay zero
            m = fmax(m, fabs(xv[j]));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        int e = 0;
        if (m > 0.0 && m < 1e300) frexp(m, &e);
        if (hl == 0) sl.scl[(int64_t)blockIdx.y * sl.Rpad + rr] = e;
        long long Xq[4];
        drp_quant4(xv, m, e, NS, Xq);
        const int chunk = ((hl >> 2) ^ (rr & 7)) << 4;           // 16-byte chunk of the row after the 128B swizzle
        for (int t = 0; t < NS; ++t)
            *reinterpret_cast<uint32_t*>(dst + ((int64_t)t * sl.Rpad + rr) * 128 + chunk + ((4 * hl) & 15)) = drp_pick4(Xq, NS - 1 - t);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// syrk: C[r][c] -= sum_{k in [k0, k0+K)} P[r][k] P[c][k] for r in [c1, R), c in [c1, min(R, col_limit)), c <= r.
// 64x64 output tile per CTA, 4 warps (2x2) of 32x32, fp64 mma.sync m8n8k4, K consumed in chunks of 64.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void syrk_tile(double* __restrict__ A, int64_t ld, int k0, int K, int c1, int R, int col_limit,
                                          int ti, int tj, double* sm, bool sync_first)
{
    double* As = sm;
    double* Bs = (ti == tj) ? As : sm + TS * SLD;
    const int tid = threadIdx.x;
    const int r0 = c1 + ti * TS, q0 = c1 + tj * TS;
    const int warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const double* ap = As + (wm * 32 + gid) * SLD + tig;
    const double* bp = Bs + (wn * 32 + gid) * SLD + tig;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kc = 0; kc < K; kc += NB) {
        const int kb = min(NB, K - kc);
        if (kc || sync_first) __syncthreads();
        for (int idx = tid; idx < TS * NB; idx += 128) {
            const int r = idx / NB, k = idx - r * NB;
            As[r * SLD + k] = (r0 + r < R && k < kb) ? A[(int64_t)(r0 + r) * ld + k0 + kc + k] : 0.0;
        }
        if (ti != tj) {
            for (int idx = tid; idx < TS * NB; idx += 128) {
                const int r = idx / NB, k = idx - r * NB;
                Bs[r * SLD + k] = (q0 + r < R && k < kb) ? A[(int64_t)(q0 + r) * ld + k0 + kc + k] : 0.0;
            }
        }
        __syncthreads();
        const int kmax = (kb + 3) & ~3;
        for (int k = 0; k < kmax; k += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ap[i * 8 * SLD + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bp[j * 8 * SLD + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    const int cmax = min(R, col_limit);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + wm * 32 + i * 8 + gid;
        if (row >= R) continue;
        double* crow = A + (int64_t)row * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = q0 + wn * 32 + j * 8 + 2 * tig;
            if (col <= row && col < cmax) crow[col] -= acc[i][j][0];
            if (col + 1 <= row && col + 1 < cmax) crow[col + 1] -= acc[i][j][1];
        }
    }
}

__global__ void __launch_bounds__(128) syrk_kernel(double* __restrict__ Mb, int64_t ld, int64_t bs, int k0, int K, int c1,
                                                   int R, int col_limit)
{
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj > ti) return;
    extern __shared__ double sm[];
    syrk_tile(Mb + (int64_t)blockIdx.z * bs, ld, k0, K, c1, R, col_limit, ti, tj, sm, false);
}

// The same update for the matrices the precision gate (factor.cuh) kept off the tensor path: a small persistent grid
// whose CTAs find an EMPTY list in the normal case and exit at once (a (Tc, T) grid of 69 KB CTAs that only exit costs
// ~75 us per launch, 2.4 ms per 2,000-leaf step).  Work item w -> (matrix of the fp64 list, tile row, tile column).
__global__ void __launch_bounds__(128) syrk_gated_kernel(double* __restrict__ Mb, int64_t ld, int64_t bs, int k0, int K, int c1,
                                                         int R, int col_limit, const int* __restrict__ zmap, int z, int T,
                                                         int Tc)
{
    const int nq = zmap[1];
    if (nq == 0) return;
    extern __shared__ double sm[];
    const int tri = Tc * (Tc + 1) / 2, pairs = tri + (T - Tc) * Tc;
    bool later = false;
    for (int w = blockIdx.x; w < nq * pairs; w += gridDim.x) {
        const int q = w / pairs;
        int v = w - q * pairs, ti, tj;
        if (v < tri) {                                 // triangular part: rows 0..Tc-1 with ti + 1 tiles each
            ti = (int)((sqrt(8.0 * v + 1.0) - 1.0) * 0.5);
            while (ti * (ti + 1) / 2 > v) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= v) ++ti;
            tj = v - ti * (ti + 1) / 2;
        } else {                                       // rows below: Tc tiles each
            v -= tri;
            ti = Tc + v / Tc;
            tj = v - (ti - Tc) * Tc;
        }
        syrk_tile(Mb + (int64_t)zmap[2 + z + q] * bs, ld, k0, K, c1, R, col_limit, ti, tj, sm, later);
        later = true;
    }
}

// Precision gate (factor.cuh): one thread sorts the z matrices into the tensor-path list and the fp64 list.
__global__ void gate_kernel(const double* __restrict__ cond, const double* __restrict__ sf, const double* __restrict__ sn,
                            double nnorm, int z, double limit, int* __restrict__ zmap)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int nt = 0, nf = 0;
    for (int i = 0; i < z; ++i) {
        double c;
        if (cond) {
            c = cond[i];
        } else {
            const double f2 = sf[i] * sf[i], n2 = sn[i] * sn[i];
            c = (nnorm * f2 + n2) / n2;
        }
        if (c <= limit) zmap[2 + nt++] = i;          // NaN / inf bounds fail the comparison -> fp64
        else zmap[2 + z + nf++] = i;
    }
    zmap[0] = nt;
    zmap[1] = nf;
}


// ---------------------------------------------------------------------------------------------------------------
// Fused panel kernels (round 2): one panel of W <= 256 columns in TWO launches instead of the chain
// potf2 -> trsm -> in-panel update -> potf2 -> trsm ... of the 64-column steps.
//
//   potrf_diag_kernel   one CTA per matrix factors the whole W x W diagonal block: per 64-column block the register
//                       potf2, then every remaining row of the diagonal region is solved by substitution (thread per
//                       row) WHILE 64 other threads run the same substitution on the identity, i.e. invert the factored
//                       64 x 64 block; the in-region trailing blocks are updated on the fp64 tensor pipe.  The block
//                       inverses Y_jj go to the workspace.
//   panel_solve_kernel  the rows below the diagonal region are independent of each other once the diagonal region is
//                       factored: a CTA owns 64 rows and runs the whole blocked forward substitution for them,
//                       X_j = (B_j - sum_{i<j} X_i L_ji^T) Y_jj^T, as 64 x 64 x 64 products on the fp64 tensor pipe.
//                       Multiplying by the explicit inverse leaves a residual of ~cond(L_jj) eps, harmless for the
//                       matrices that take the split-integer update anyway (cond <= 3.4e5, backward error 2e-11); every
//                       other matrix (the precision gate's fp64 list, callers without a bound) gets one step of
//                       fixed-precision refinement, X += (B - X L_jj^T) Y_jj^T, which restores the backward error of a
//                       substitution (Skeel).
// ---------------------------------------------------------------------------------------------------------------
constexpr int GLD = 68;            // row stride of the 64 x 64 operand tiles (conflict-free mma fragments, 16-byte rows)
constexpr int GT = NB * GLD;       // doubles per tile

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// 64 x 64 tile <- global rows (asynchronously): entry (r, c) = src[r * ld + c] for r < nr, c < nc (and c <= r if `lower`),
// otherwise `ident` ? (r == c) : 0.  All 256 threads; coalesced 8-byte copies.
__device__ __forceinline__ void tile_load_async(double* dst, const double* src, int64_t ld, int nr, int nc, bool lower, bool ident)
{
    const int c = threadIdx.x & 63, rb = threadIdx.x >> 6;
#pragma unroll 4
    for (int r = rb; r < NB; r += 4) {
        if (r < nr && c < nc && (!lower || c <= r)) cp_async8(dst + r * GLD + c, src + (int64_t)r * ld + c);
        else dst[r * GLD + c] = (ident && r == c) ? 1.0 : 0.0;
    }
}

// acc += As[rows 16 wm ..][k] * Bs[cols 32 wn ..][k] over k < 64 ("NT" product of two row-major tiles); 8 warps = 4 x 2.
// acc[i][j][0/1] = entry (16 wm + 8 i + gid, 32 wn + 8 j + 2 tig + 0/1).
// LOWER: Bs is lower triangular (a diagonal block of L or of its inverse), so output columns [8 q, 8 q + 8) only see
// k < 8 q + 8 - 44 % of the tensor instructions of a full product are skipped (warp-uniform predicates).
template <bool LOWER>
__device__ __forceinline__ void blk_gemm_nt(double (&acc)[2][4][2], const double* As, const double* Bs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int wn = warp & 1;
    const double* ap = As + ((warp >> 1) * 16 + gid) * GLD + tig;
    const double* bp = Bs + (wn * 32 + gid) * GLD + tig;
    const int kend = LOWER ? wn * 32 + 32 : NB;
#pragma unroll 4
    for (int k = 0; k < kend; k += 4) {
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = ap[i * 8 * GLD + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = bp[j * 8 * GLD + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (LOWER && k >= wn * 32 + 8 * j + 8) continue;
#pragma unroll
            for (int i = 0; i < 2; ++i) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
}
__device__ __forceinline__ void acc_zero(double (&acc)[2][4][2])
{
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// x <- x L^-T for one row in registers: L (lower, row stride GLD) and invd = 1 / diag(L) in shared memory (broadcast reads).
__device__ __forceinline__ void row_solve(double (&x)[NB], const double* Ls, const double* invd)
{
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const double* lc = Ls + c * GLD;
        // EIGHT accumulation chains: in potrf_diag_kernel this runs on one or two warps per SM sub-partition, a pure
        // dependent-latency chain of 64 stages
        double a0 = x[c], a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
#pragma unroll
        for (int p = 0; p + 7 < c; p += 8) {
            const double2 l01 = *reinterpret_cast<const double2*>(lc + p);
            const double2 l23 = *reinterpret_cast<const double2*>(lc + p + 2);
            const double2 l45 = *reinterpret_cast<const double2*>(lc + p + 4);
            const double2 l67 = *reinterpret_cast<const double2*>(lc + p + 6);
            a0 = fma(-x[p], l01.x, a0);
            a1 = fma(-x[p + 1], l01.y, a1);
            a2 = fma(-x[p + 2], l23.x, a2);
            a3 = fma(-x[p + 3], l23.y, a3);
            a4 = fma(-x[p + 4], l45.x, a4);
            a5 = fma(-x[p + 5], l45.y, a5);
            a6 = fma(-x[p + 6], l67.x, a6);
            a7 = fma(-x[p + 7], l67.y, a7);
        }
#pragma unroll
        for (int p = c & ~7; p < c; ++p) {
            if ((p & 1) == 0) a0 = fma(-x[p], lc[p], a0);
            else a1 = fma(-x[p], lc[p], a1);
        }
        x[c] = (((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7))) * invd[c];
    }
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* Mb, int64_t ld, int64_t bs, int c0, int W,
                                                         int32_t* __restrict__ info, double* __restrict__ linv)
{
    extern __shared__ __align__(16) double sm[];
    double* Ld = sm;                   // factored diagonal block: zeros above the diagonal, identity beyond its nb columns
    double* Yd = Ld + GT;              // its inverse
    double* T = Yd + GT;               // up to three tiles: rows of the diagonal region below the block (block column j)
    __shared__ __align__(16) double Lp[LP_ELEMS];
    __shared__ double invd[NB];
    __shared__ int sbad;
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15;
    double* A = Mb + (int64_t)blockIdx.x * bs + (int64_t)c0 * ld + c0;
    double* Yg = linv + (int64_t)blockIdx.x * (4 * NB * NB);
    const int nblk = (W + NB - 1) / NB;
    for (int j = 0; j < nblk; ++j) {
        const int nbj = min(NB, W - NB * j);
        const int nbelow = W - NB * (j + 1);                        // rows of the region below this block (<= 192)
        double* Ajj = A + (int64_t)(NB * j) * ld + NB * j;
        // rows below, block column j: on their way to shared memory while the diagonal block is factored
        for (int t = 0; t * NB < nbelow; ++t)
            tile_load_async(T + t * GT, Ajj + (int64_t)(NB * (t + 1)) * ld, ld, nbelow - t * NB, NB, false, false);
        if (tid == 0) sbad = 0;
        {
            double a[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = 4 * ty + r, q = 4 * tx + k;
                    a[r][k] = (i < nbj && q <= i) ? Ajj[(int64_t)i * ld + q] : (i == q ? 1.0 : 0.0);
                }
            __syncthreads();
            potf2_regs(a, Lp, &sbad);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = 4 * ty + r, q = 4 * tx + k;
                    if (i < nbj && q <= i) Ajj[(int64_t)i * ld + q] = a[r][k];
                    Ld[i * GLD + q] = q <= i ? a[r][k] : 0.0;
                }
        }
        cp_async_wait_all();
        __syncthreads();
        if (tid == 0 && sbad && sbad <= nbj && info) atomicCAS(info + blockIdx.x, 0, c0 + NB * j + sbad);
        if (tid < NB) invd[tid] = 1.0 / Ld[tid * GLD + tid];
        __syncthreads();
        {   // threads 0..63: row t of the identity -> column t of Y = L^-1; threads 64..: one row of the region below
            const int rr = tid - NB;
            const bool inv_row = tid < NB;
            if (inv_row || rr < nbelow) {
                double x[NB];
                double* trow = T + (rr >> 6) * GT + (rr & 63) * GLD;
                if (inv_row) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) x[c] = (c == tid) ? 1.0 : 0.0;
                } else {
#pragma unroll
                    for (int c = 0; c < NB; c += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(trow + c);
                        x[c] = v.x;
                        x[c + 1] = v.y;
                    }
                }
                row_solve(x, Ld, invd);
                if (inv_row) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) Yd[c * GLD + tid] = x[c];
                } else {
#pragma unroll
                    for (int c = 0; c < NB; c += 2) *reinterpret_cast<double2*>(trow + c) = make_double2(x[c], x[c + 1]);
                }
            }
        }
        __syncthreads();
        {   // the inverse to the workspace, the solved rows back to the matrix (coalesced)
            const int c = tid & 63, rb = tid >> 6;
            for (int r = rb; r < NB; r += 4) Yg[(j * NB + r) * NB + c] = Yd[r * GLD + c];
            for (int r = rb; r < nbelow; r += 4)
                Ajj[(int64_t)(NB + r) * ld + c] = T[(r >> 6) * GT + (r & 63) * GLD + c];
        }
        // trailing blocks of the diagonal region: A_ik -= A_ij A_kj^T for j < k <= i
        const int lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3, wm = warp >> 1, wn = warp & 1;
        for (int i = j + 1; i < nblk; ++i)
            for (int k = j + 1; k <= i; ++k) {
                if (i == k && 32 * wn > 16 * wm + 15) continue;               // this warp's window lies above the diagonal
                double acc[2][4][2];
                acc_zero(acc);
                blk_gemm_nt<false>(acc, T + (i - j - 1) * GT, T + (k - j - 1) * GT);
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    const int row = NB * i + wm * 16 + ii * 8 + gid;
                    if (row >= W) continue;
                    double* crow = A + (int64_t)row * ld;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int col = NB * k + wn * 32 + jj * 8 + 2 * tig;
                        if (col <= row && col < W) crow[col] -= acc[ii][jj][0];
                        if (col + 1 <= row && col + 1 < W) crow[col + 1] -= acc[ii][jj][1];
                    }
                }
            }
        __syncthreads();               // the block's global writes are visible to its own later reads
    }
}

__global__ void __launch_bounds__(256) panel_solve_kernel(double* Mb, int64_t ld, int64_t bs, int c0, int W, int R,
                                                          const double* __restrict__ linv, const int* __restrict__ zmap,
                                                          int z, int refine_all, const TrsmSlice sl)
{
    extern __shared__ __align__(16) double sm[];
    const int nblk = (W + NB - 1) / NB;
    double* X = sm;                    // nblk tiles: this CTA's 64 rows of the panel
    double* Lt = sm + nblk * GT;       // the block of L (or of the inverse) being multiplied
    const int zz = blockIdx.y, tid = threadIdx.x;
    double* A = Mb + (int64_t)zz * bs;
    const double* Yg = linv + (int64_t)zz * (4 * NB * NB);
    const int r0 = c0 + W + blockIdx.x * NB;
    const int nrows = max(0, min(NB, R - r0));     // 0: a CTA past the last row, there to write the zero rows of the slice buffer
    bool refine = refine_all != 0;
    if (!refine && zmap) {             // on the precision gate's fp64 list?
        const int nf = zmap[1];
        for (int q = 0; q < nf; ++q) refine |= (zmap[2 + z + q] == zz);
    }
    for (int j = 0; j < nblk; ++j)
        tile_load_async(X + j * GT, A + (int64_t)r0 * ld + c0 + NB * j, ld, nrows, W - NB * j, false, false);
    const int lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3, wm = warp >> 1, wn = warp & 1;
    for (int j = 0; j < (nrows > 0 ? nblk : 0); ++j) {
        const int nbj = min(NB, W - NB * j);
        double* Xj = X + j * GT;
        const double* Ljrow = A + (int64_t)(c0 + NB * j) * ld + c0;            // row block j of the factored region
        double acc[2][4][2];
        if (j > 0) {
            acc_zero(acc);
            for (int i = 0; i < j; ++i) {
                __syncthreads();                                               // Lt free
                tile_load_async(Lt, Ljrow + NB * i, ld, nbj, NB, false, false);
                cp_async_wait_all();
                __syncthreads();
                blk_gemm_nt<false>(acc, X + i * GT, Lt);
            }
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    double2* p = reinterpret_cast<double2*>(Xj + (wm * 16 + ii * 8 + gid) * GLD + wn * 32 + jj * 8 + 2 * tig);
                    double2 v = *p;
                    v.x -= acc[ii][jj][0];
                    v.y -= acc[ii][jj][1];
                    *p = v;
                }
        }
        __syncthreads();
        tile_load_async(Lt, Yg + j * NB * NB, NB, NB, NB, false, false);
        cp_async_wait_all();
        __syncthreads();
        acc_zero(acc);
        blk_gemm_nt<true>(acc, Xj, Lt);                                              // X0 = B Y^T
        if (refine) {
            double b[2][4][2], res[2][4][2];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const double2 v = *reinterpret_cast<const double2*>(Xj + (wm * 16 + ii * 8 + gid) * GLD + wn * 32 + jj * 8 + 2 * tig);
                    b[ii][jj][0] = v.x;
                    b[ii][jj][1] = v.y;
                }
            __syncthreads();                                                   // every read of B and of Y is done
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<double2*>(Xj + (wm * 16 + ii * 8 + gid) * GLD + wn * 32 + jj * 8 + 2 * tig) =
                        make_double2(acc[ii][jj][0], acc[ii][jj][1]);
            tile_load_async(Lt, Ljrow + NB * j, ld, nbj, nbj, true, true);     // L_jj
            cp_async_wait_all();
            __syncthreads();
            acc_zero(res);
            blk_gemm_nt<true>(res, Xj, Lt);                                          // X0 L^T
            __syncthreads();
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<double2*>(Xj + (wm * 16 + ii * 8 + gid) * GLD + wn * 32 + jj * 8 + 2 * tig) =
                        make_double2(b[ii][jj][0] - res[ii][jj][0], b[ii][jj][1] - res[ii][jj][1]);
            tile_load_async(Lt, Yg + j * NB * NB, NB, NB, NB, false, false);
            cp_async_wait_all();
            __syncthreads();
            blk_gemm_nt<true>(acc, Xj, Lt);                                          // X0 + (B - X0 L^T) Y^T
        }
        __syncthreads();                                                       // every read of the tile is done
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int r = wm * 16 + ii * 8 + gid;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = wn * 32 + jj * 8 + 2 * tig;
                *reinterpret_cast<double2*>(Xj + r * GLD + c) = make_double2(acc[ii][jj][0], acc[ii][jj][1]);
                if (r < nrows) {
                    double* g = A + (int64_t)(r0 + r) * ld + c0 + NB * j + c;
                    if (c < nbj) g[0] = acc[ii][jj][0];
                    if (c + 1 < nbj) g[1] = acc[ii][jj][1];
                }
            }
        }
    }
    if (sl.sl == nullptr) return;
    // ---- the solved rows ARE the panel of the trailing update that follows (all W columns on chip, so the row scale is
    //      known): its int8 slices go out from shared memory in the layout of oz_slice_kernel (ozaki.cu), bit-identical
    cp_async_wait_all();
    __syncthreads();
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) sl.counter[0] = sl.counter[1] = 0;
    const int NS = sl.S, KBN = sl.KBN;
    uint8_t* dst = sl.sl + (int64_t)zz * sl.sl_bs;
    for (int it = 0; it < NB / 8; ++it) {
        const int i = warp * (NB / 8) + it;                        // row inside the CTA
        const int rr = blockIdx.x * NB + i;                        // row of the slice buffer (relative to c0 + W)
        double xv[8];
        double m = 0.0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const int tile = 2 * kb + (lane >> 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double v = tile < nblk ? X[tile * GT + i * GLD + 4 * (lane & 15) + q] : 0.0;   // rows >= nrows are zeros
                xv[kb * 4 + q] = v;
                m = fmax(m, fabs(v));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        int e = 0;
        if (m > 0.0 && m < 1e300) frexp(m, &e);
        if (lane == 0) sl.scl[(int64_t)zz * sl.Rpad + rr] = e;
        const int chunk = ((lane >> 2) ^ (rr & 7)) << 4;           // 16-byte chunk of the row after the 128B swizzle
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            if (kb >= KBN) break;
            long long Xq[4];
            const double xk[4] = {xv[kb * 4], xv[kb * 4 + 1], xv[kb * 4 + 2], xv[kb * 4 + 3]};
            drp_quant4(xk, m, e, NS, Xq);
            for (int t = 0; t < NS; ++t)
                *reinterpret_cast<uint32_t*>(dst + ((int64_t)(t * KBN + kb) * sl.Rpad + rr) * 128 + chunk + ((4 * lane) & 15)) =
                    drp_pick4(Xq, NS - 1 - t);
        }
    }
}

}  // namespace

#include <stdlib.h>
static const int TRSM_SMEM = (NB * NB + NB + PR * PLD) * (int)sizeof(double);
static const int SYRK_SMEM = 2 * TS * SLD * (int)sizeof(double);

void drp_dmma_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit, const int* zmap,
                   cudaStream_t st)
{
    if (drp_first_on_device(0)) {
        cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM);
        cudaFuncSetAttribute(syrk_gated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM);
    }
    const int below = R - c1;
    const int cl = min(R, col_limit) - c1;
    if (below <= 0 || cl <= 0) return;
    const int T = (below + TS - 1) / TS;
    const int Tc = min(T, (cl + TS - 1) / TS);
    // algorithmic flops: 2*K per updated entry of the lower trapezoid rows [c1,R) x cols [c1,c1+cl)
    const double entries = (double)cl * (cl + 1) / 2.0 + (double)(below - cl) * cl;
    if (zmap) {                 // gated: normally nothing to do (work is not attributed; the tensor-path launch carries it)
        DRP_LAUNCH(DRP_K_SYRK, 0.0, st);
        syrk_gated_kernel<<<296, 128, SYRK_SMEM, st>>>(M, ld, bs, k0, K, c1, R, col_limit, zmap, z, T, Tc);
        return;
    }
    DRP_LAUNCH(DRP_K_SYRK, (double)z * entries * 2.0 * K, st);
    syrk_kernel<<<dim3(Tc, T, z), 128, SYRK_SMEM, st>>>(M, ld, bs, k0, K, c1, R, col_limit);
}

// Trailing update of rows [c1, R), columns [col_lo, min(R, col_hi)): int8 tcgen05 pipe when a workspace is given and the
// (whole) update is large, fp64 mma.sync otherwise.  do_slice / buffer / counter_idx: see drp_ozaki_syrk.
static int trailing_update(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_lo, int col_hi,
                           void* ws, int64_t ws_bytes, const int* zmap, int do_slice, int buffer, int counter_idx,
                           cudaStream_t st)
{
    const int rc = drp_ozaki_syrk(M, ld, bs, z, k0, K, c1, R, col_lo, col_hi, drp_ozaki_slices(), ws, ws_bytes, zmap, do_slice,
                                  buffer, counter_idx, st);
    if (rc > 1 || rc < 0) return rc;
    // rc == 1: the shape does not qualify for the tensor path -> every matrix on the fp64 kernel;
    // rc == 0: the matrices the precision gate kept off the tensor path (normally none: the CTAs find an empty list)
    drp_dmma_syrk(M, ld, bs, z, k0, K, col_lo > c1 ? col_lo : c1, R, col_hi, rc == 1 ? nullptr : zmap, st);
    return 0;
}

// Would the tensor path take an update of this shape?  (Look-ahead only splits updates that do.)
static bool tensor_update_qualifies(int K, int c1, int R, int col_limit)
{
    const int cmax = R < col_limit ? R : col_limit;
    const int KBN = K > 128 ? 2 : 1;
    return K >= 32 && K <= 256 && drp_ozaki_slices() * KBN <= 12 && R - c1 >= 384 && cmax - c1 >= 256;
}

static int outer_block(int R_full, bool tensor_path)
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_OUTER_BLOCK");
        v = e ? atoi(e) : 0;
        if (v != 0 && v < NB) v = NB;
        v = (v / NB) * NB;
    }
    if (v) return v;
    return (tensor_path && R_full >= 1024) ? 256 : 128;
}

static bool fused_slicing()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_FUSED_SLICING");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

// Fused panel kernels (potrf_diag + panel_solve) instead of the 64-column chain: DRP_PANEL_WIDTH = 0 (off: potf2 / trsm
// steps), 64, 128 (default) or 256 columns per fused panel.  They need the workspace (block inverses).
static int panel_width()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_PANEL_WIDTH");
        v = e ? atoi(e) : 128;
        if (v != 0 && v != 64 && v != 128 && v != 256) v = 128;
    }
    return v;
}

// Slices of the next update written by panel_solve_kernel itself (instead of oz_slice_kernel re-reading the solved rows).
// Parity-identical; measured at 2,000 leaves on one B200 the slicing work moves from one kernel into the other (slicing
// 0.78 -> 0.45 ms, panel solve 1.27 -> 1.63 ms per step) and the replayed step is 0.1 ms SLOWER (21.47 vs 21.35 ms), so it
// is opt-in (DRP_PANEL_SLICING=1).
static bool panel_slicing()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_PANEL_SLICING");
        v = e ? atoi(e) : 0;
    }
    return v != 0;
}

static bool lookahead_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DRP_LOOKAHEAD");
        v = e ? atoi(e) : 0;        // off by default: measured, it does not pay on one GPU (see the comment at drp_factor_run)
    }
    return v != 0;
}

// Two-level right-looking elimination: panels of `outer_block()` columns are factored with NB-wide inner steps whose
// trailing updates stay inside the panel; the rest of the matrix is then updated once per panel with K = panel width
// (halves / quarters the read-modify-write traffic of the trailing matrix compared with K = NB, and gives the tensor
// path a contraction length of 128-256).
//
// Look-ahead (aux != null).  The factorisation of an outer panel is a latency chain of ~20 narrow kernels (4 x {potf2, trsm,
// slicing, in-panel update}) that leaves the GPU idle; the update that follows is throughput work.  They only depend on
// each other through the columns of the NEXT panel, so the update of panel k is split by columns: part (a), the next
// panel's columns, runs first on the caller's stream; part (b), everything right of them, goes to the auxiliary stream and
// runs WHILE panel k+1 is being factored (which reads and writes the next panel's columns only).  Update k+1 waits for
// part (b) of update k.  Part (b) re-uses the slices part (a) made (buffer 0); the in-panel updates of panel k+1 slice into
// buffer 1.  Two events, one auxiliary stream, all supplied by the caller; everything is joined back before returning.
// MEASURED (round 2, profiles/r02/lookahead_ab.txt): parity-green but NO gain at 2,000 leaves on one B200 - 22.14 vs
// 22.11 ms per step, conditional posterior 10.15 vs 10.15 ms, with persistent or one-unit-per-CTA part (b) kernels and
// stream priorities either way.  Both the chain's wide kernels (trsm: 930 CTAs, slicing) and part (b) want every SM; what
// the overlap recovers (idle SMs under potf2, launch gaps) it loses again in contention (trsm 1.35 -> 3.4 ms inside the
// overlapped region).  Hence opt-in (DRP_LOOKAHEAD=1); the column-range form of the update it needs is exercised by the
// GPU test-suite either way (tests/test_gpu_ozaki.py::test_split_update_equals_whole).
int drp_factor_run(double* M, int64_t ld, int64_t bs, int z, int nfactor, int R_full, int grow_base, int col_limit,
                   int32_t* info, void* ws, int64_t ws_bytes, const DrpGate& gate, const DrpAux* aux, cudaStream_t st)
{
    if (drp_first_on_device(1)) cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM);
    if (drp_first_on_device(2)) {
        cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * GT * (int)sizeof(double));
        cudaFuncSetAttribute(panel_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * GT * (int)sizeof(double));
    }
    if (info) cudaMemsetAsync(info, 0, sizeof(int32_t) * z, st);
    const bool gated = gate.cond != nullptr || (gate.sf != nullptr && gate.sn != nullptr);
    const bool tensor_path = gated && ws != nullptr && drp_ozaki_slices() != 0 && ws_bytes >= drp_ozaki_workspace_bytes(R_full, z)
                             && R_full >= 384 + NB;                 // smaller matrices never reach a qualifying update
    // the block inverses of the fused panel kernels live at the very end of the workspace
    double* linv = nullptr;
    if (ws != nullptr && panel_width() != 0 && ws_bytes >= drp_linv_bytes(z) + drp_zmap_bytes(z)) {
        ws_bytes -= drp_linv_bytes(z);
        linv = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + ws_bytes + 15) & ~(uintptr_t)15);
    }
    const int NBO = outer_block(R_full, tensor_path);
    void* tws = tensor_path ? ws : nullptr;
    int* zmap = nullptr;
    if (tensor_path) {
        ws_bytes -= drp_zmap_bytes(z);
        zmap = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ws_bytes);
        DRP_LAUNCH(DRP_K_OTHER, 16.0 * z, st);
        gate_kernel<<<1, 32, 0, st>>>(gate.cond, gate.sf, gate.sn, gate.nnorm, z, drp_ozaki_cond_limit(), zmap);
    }
    const bool look = tensor_path && aux != nullptr && aux->stream != nullptr && aux->ev_fork != nullptr &&
                      aux->ev_join != nullptr && lookahead_enabled();
    bool pending_b = false;                                         // a part (b) is in flight on the auxiliary stream
    int rc = 0;
    for (int C0 = 0; C0 < nfactor && rc == 0; C0 += NBO) {
        const int W = min(NBO, nfactor - C0);
        const int C1 = C0 + W;
        const int R = grow_base > 0 ? min(R_full, grow_base + C1) : R_full;
        // inside the outer panel: 64-wide steps; with the tensor path the panel is split again into 128-wide
        // sub-panels so that only 64-column-wide updates are left to the fp64 mma.sync kernel
        bool outer_sliced = false;
        const int NBM = linv ? min(W, panel_width()) : (tensor_path && W > 2 * NB) ? 2 * NB : W;
        for (int Q0 = C0; Q0 < C1 && rc == 0; Q0 += NBM) {
            const int Q1 = min(Q0 + NBM, C1);
            bool inner_sliced = false;
            if (linv) {                // the whole sub-panel in two launches (fused panel kernels)
                const int Wq = Q1 - Q0, nblk = (Wq + NB - 1) / NB;
                {
                    DRP_LAUNCH(DRP_K_POTRF_DIAG, (double)z * Wq * Wq * Wq / 3.0, st);
                    potrf_diag_kernel<<<z, 256, (2 + (nblk > 1 ? nblk - 1 : 1)) * GT * sizeof(double), st>>>(M, ld, bs, Q0, Wq, info, linv);
                }
                const int below = R - Q1;
                if (below > 0) {
                    // the trailing update that consumes exactly this launch's columns takes its slices from it: the update
                    // inside the outer panel (Q1 < C1), or the update of the rest of the matrix when the outer panel is one piece
                    TrsmSlice psl{nullptr, 0, nullptr, nullptr, 0, 0, 0};
                    int gx = (below + NB - 1) / NB;
                    if (panel_slicing() && tws != nullptr && !look && (Q1 < C1 || Q0 == C0)) {
                        DrpOzLayout lay;
                        const bool inner = Q1 < C1;
                        if (drp_ozaki_layout(z, Wq, Q1, R, inner ? min(col_limit, C1) : col_limit, drp_ozaki_slices(), tws, ws_bytes,
                                             inner ? 1 : 0, &lay) && lay.rb == Q1 && lay.Rpad >= below) {
                            psl = TrsmSlice{lay.sl, lay.sl_bs, lay.scl, lay.counter, lay.Rpad, drp_ozaki_slices(), lay.KBN};
                            gx = lay.Rpad / NB;
                            (inner ? inner_sliced : outer_sliced) = true;
                        }
                    }
                    DRP_LAUNCH(DRP_K_PANEL_SOLVE, (double)z * below * Wq * Wq, st);
                    panel_solve_kernel<<<dim3(gx, z), 256, (nblk + 1) * GT * sizeof(double), st>>>(
                        M, ld, bs, Q0, Wq, R, linv, zmap, z, tensor_path ? 0 : 1, psl);
                }
            }
            for (int c0 = Q0; c0 < Q1 && rc == 0 && !linv; c0 += NB) {
                const int nb = min(NB, Q1 - c0);
                const int c1 = c0 + nb;
                {
                    DRP_LAUNCH(DRP_K_POTF2, (double)z * nb * nb * nb / 3.0, st);
                    potf2_kernel<<<z, 256, 0, st>>>(M, ld, bs, c0, nb, info);
                }
                const int below = R - c1;
                if (below > 0) {
                    dim3 g((below + PR - 1) / PR, z);
                    // the in-panel update that follows a full 64-column step takes its slices from the solve itself
                    TrsmSlice tsl{nullptr, 0, nullptr, nullptr, 0, 0};
                    DrpOzLayout lay;
                    const bool fused = fused_slicing() && tws != nullptr && c1 < Q1 && nb == NB && (c1 & 7) == 0 &&
                                       drp_ozaki_layout(z, nb, c1, R, min(col_limit, Q1), drp_ozaki_slices(), tws, ws_bytes, 1, &lay) &&
                                       lay.Rpad == (int)g.x * PR;
                    if (fused) tsl = TrsmSlice{lay.sl, lay.sl_bs, lay.scl, lay.counter, lay.Rpad, drp_ozaki_slices()};
                    {
                        DRP_LAUNCH(DRP_K_TRSM, (double)z * below * nb * nb, st);
                        trsm_kernel<<<g, PR, TRSM_SMEM, st>>>(M, ld, bs, c0, nb, R, tsl);
                    }
                    if (c1 < Q1)                                                                // inside the sub-panel
                        rc = trailing_update(M, ld, bs, z, c0, nb, c1, R, c1, min(col_limit, Q1), tws, ws_bytes, zmap, fused ? 0 : 1, 1,
                                             0, st);
                }
            }
            if (Q1 < C1 && rc == 0)                                                                     // rest of the outer panel
                rc = trailing_update(M, ld, bs, z, Q0, Q1 - Q0, Q1, R, Q1, min(col_limit, C1), tws, ws_bytes, zmap, inner_sliced ? 0 : 1,
                                     1, 0, st);
        }
        if (rc) break;
        // ---- rest of the matrix
        if (pending_b) {                                            // part (b) of the previous update writes the same entries
            cudaStreamWaitEvent(st, aux->ev_join, 0);
            pending_b = false;
        }
        const int Wn = min(NBO, nfactor - C1);                      // width of the next outer panel (0: this was the last)
        const int cmax = min(R, col_limit);
        if (look && Wn >= NB && C1 + Wn < cmax && tensor_update_qualifies(W, C1, R, col_limit)) {
            rc = trailing_update(M, ld, bs, z, C0, W, C1, R, C1, C1 + Wn, tws, ws_bytes, zmap, 1, 0, 0, st);          // (a)
            if (rc) break;
            cudaEventRecord(aux->ev_fork, st);
            cudaStreamWaitEvent(aux->stream, aux->ev_fork, 0);
            rc = trailing_update(M, ld, bs, z, C0, W, C1, R, C1 + Wn, col_limit, tws, ws_bytes, zmap, 0, 0, 1, aux->stream);   // (b)
            cudaEventRecord(aux->ev_join, aux->stream);
            pending_b = true;
        } else {
            rc = trailing_update(M, ld, bs, z, C0, W, C1, R, C1, col_limit, tws, ws_bytes, zmap, outer_sliced ? 0 : 1, 0, 0, st);
        }
    }
    if (pending_b) cudaStreamWaitEvent(st, aux->ev_join, 0);        // the auxiliary stream is joined before anything follows
    if (rc) return rc;
    return drp_launch_status();
}

// Plain batched Cholesky of z dense n x n matrices (lower), optionally zeroing the strict upper triangle so the
// result equals torch.linalg.cholesky's.  Replaces the call inside MultivariateNormal.__init__ (S/models.py:118).
namespace {
__global__ void zero_upper_kernel(double* __restrict__ A, int n, int64_t ld, int64_t bs)
{
    const int r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && c > r) A[(int64_t)blockIdx.z * bs + (int64_t)r * ld + c] = 0.0;
}
}  // namespace

DRP_API int64_t drp_factor_workspace_bytes(int rows, int z) { return drp_ozaki_workspace_bytes(rows, z); }
DRP_API double drp_ozaki_cond_limit_f64(void) { return drp_ozaki_cond_limit(); }

// Test hook: one trailing update C[r][c] -= sum_{k in [k0,k0+K)} P[r][k] P[c][k] (rows [c1,R), cols [c1,min(R,col_limit)),
// c <= r) with slices = 0: fp64 mma.sync kernel, 4..7: int8 tcgen05 kernel with that many slices.  Returns 1 if the
// tcgen05 kernel does not accept the shape.
DRP_API int drp_syrk_update_f64(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit,
                                int slices, void* ws, int64_t ws_bytes, void* stream)
{
    if (!M) return -1;
    if (z <= 0 || K <= 0 || R <= c1) return -4;
    cudaStream_t st = (cudaStream_t)stream;
    if (slices == 0) {
        drp_dmma_syrk(M, ld, bs, z, k0, K, c1, R, col_limit, nullptr, st);
        return drp_launch_status();
    }
    return drp_ozaki_syrk(M, ld, bs, z, k0, K, c1, R, c1, col_limit, slices, ws, ws_bytes, nullptr, 1, 0, 0, st);
}

// Test hook: the column range [col_lo, col_hi) of the same update (rows [c1, R), c <= r); do_slice = 0 re-uses the slices of
// the preceding call with the same panel, counter_idx selects the work-unit counter (two parts may run on two streams).
DRP_API int drp_syrk_update_range_f64(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_lo,
                                      int col_hi, int slices, void* ws, int64_t ws_bytes, int do_slice, int counter_idx,
                                      void* stream)
{
    if (!M) return -1;
    if (z <= 0 || K <= 0 || R <= c1 || col_lo < c1) return -4;
    cudaStream_t st = (cudaStream_t)stream;
    if (slices == 0) {
        drp_dmma_syrk(M, ld, bs, z, k0, K, col_lo, R, col_hi, nullptr, st);
        return drp_launch_status();
    }
    return drp_ozaki_syrk(M, ld, bs, z, k0, K, c1, R, col_lo, col_hi, slices, ws, ws_bytes, nullptr, do_slice, 0, counter_idx, st);
}

DRP_API int drp_potrf_batched_f64(double* A, int n, int z, int zero_upper, int32_t* info, const double* cond_bound, void* ws,
                                  int64_t ws_bytes, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!A) return -1;
    if (n <= 0) return -2;
    if (z <= 0) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    const DrpGate gate{cond_bound, nullptr, nullptr, 0.0};
    int rc = drp_factor_run(A, n, (int64_t)n * n, z, n, n, 0, n, info, ws, ws_bytes, gate, &aux, st);
    if (rc) return rc;
    if (zero_upper) {
        dim3 g((n + 255) / 256, n, z);
        DRP_LAUNCH(DRP_K_OTHER, (double)z * n * n * 4.0, st);
        zero_upper_kernel<<<g, 256, 0, st>>>(A, n, n, (int64_t)n * n);
    }
    return drp_launch_status();
}
