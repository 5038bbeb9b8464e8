// Ornstein-Uhlenbeck covariance assembly, multivariate-normal log-prob with analytic gradients, and the
// conditional internal|leaves posterior — all on top of the partial-Cholesky engine in factor.cu.
//
// Reference (S/ = /root/reference/draupnir/src/draupnir/):
//   OUKernel_Fast.forward                       S/models_utils.py:303-311   K_z = sf_z^2 exp(-t/lam_z) + sn_z^2 I
//   DRAUPNIRModelClass.gp_prior                 S/models.py:85-121          sum_z log N(x_z; 0, K_z)
//   conditional_sampling / conditional_samplingMAP   S/models.py:149-242    p(z_internal | z_leaves)
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"

namespace {

constexpr double LOG_2PI = 1.8378770664093454835606594728112;

// ---------------------------------------------------------------------------------------------------------------
// K2: dense OU covariance [z,n,n].  One thread per (row, col); the distance is loaded once and reused for all z
// latent dimensions, every store is a coalesced 8-byte-per-lane row segment.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ou_cov_fwd_kernel(const double* __restrict__ D, int64_t ldd, int n,
                                                         const double* __restrict__ sf, const double* __restrict__ sn,
                                                         const double* __restrict__ lam, int z, double* __restrict__ K)
{
    extern __shared__ double hp[];   // [3][z]: sf^2, sn^2, -1/lam
    for (int i = threadIdx.x; i < z; i += blockDim.x) {
        hp[i] = sf[i] * sf[i];
        hp[z + i] = sn[i] * sn[i];
        hp[2 * z + i] = -1.0 / lam[i];
    }
    __syncthreads();
    const int r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double d = D[(int64_t)r * ldd + c];
    double* out = K + (int64_t)r * n + c;
    const int64_t bs = (int64_t)n * n;
    for (int zz = 0; zz < z; ++zz) {
        double v = hp[zz] * drp_exp(d * hp[2 * z + zz]);
        if (r == c) v += hp[z + zz];
        out[zz * bs] = v;
    }
}

// Backward of K2 given G = dLoss/dK [z,n,n]:  partial sums of G.E, G.E.D and diag(G) per (z, row).
__global__ void __launch_bounds__(256) ou_cov_bwd_kernel(const double* __restrict__ D, int64_t ldd, int n,
                                                         const double* __restrict__ lam, const double* __restrict__ G,
                                                         double* __restrict__ part /*[z][n][3]*/)
{
    __shared__ double red[32];
    const int r = blockIdx.x, zz = blockIdx.y;
    const double nil = -1.0 / lam[zz];
    const double* g = G + ((int64_t)zz * n + r) * n;
    const double* drow = D + (int64_t)r * ldd;
    double sE = 0.0, sED = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const double d = drow[c];
        const double t = g[c] * drp_exp(d * nil);
        sE += t;
        sED += t * d;
    }
    sE = block_sum(sE, red);
    sED = block_sum(sED, red);
    if (threadIdx.x == 0) {
        double* p = part + ((int64_t)zz * n + r) * 3;
        p[0] = sE;
        p[1] = sED;
        p[2] = g[r];
    }
}

// Deterministic second stage: out[z][k] = sum_r part[z][r][k].
__global__ void __launch_bounds__(256) reduce3_kernel(const double* __restrict__ part, int n, double* __restrict__ out)
{
    __shared__ double red[32];
    const int zz = blockIdx.x;
    for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int r = threadIdx.x; r < n; r += blockDim.x) s += part[((int64_t)zz * n + r) * 3 + k];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[zz * 3 + k] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Assembly of the augmented matrix (see factor.cu) straight from the patristic distances.  Lower part only.
//   identity != 0 : second block = I (prior: yields -K^-1 and -alpha)
//   identity == 0 : second block = K_ab / K_aa gathered through idx[n..n+ni) (conditional posterior)
// idx (nullable) maps position -> node row of D; rows/cols 0..n-1 are the conditioning (leaf) nodes.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ou_assemble_kernel(const double* __restrict__ D, int64_t ldd,
                                                          const int32_t* __restrict__ idx, int n, int n2, int identity,
                                                          const double* __restrict__ sf, const double* __restrict__ sn,
                                                          const double* __restrict__ lam, const double* __restrict__ x,
                                                          int z, double* __restrict__ M, int64_t ld, int64_t bs)
{
    const int r = blockIdx.y;
    if ((int)(blockIdx.x * blockDim.x) > r) return;
    extern __shared__ double hp[];
    for (int i = threadIdx.x; i < z; i += blockDim.x) {
        hp[i] = sf[i] * sf[i];
        hp[z + i] = sn[i] * sn[i];
        hp[2 * z + i] = -1.0 / lam[i];
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > r) return;
    double* out = M + (int64_t)r * ld + c;
    // node behind a row/col position (-1: the x row/col)
    const int pr = r < n ? r : (r == n ? -1 : r - 1);
    const int pc = c < n ? c : (c == n ? -1 : c - 1);
    if (pr < 0 || pc < 0) {                       // right-hand-side row, or the (unused) column n
        for (int zz = 0; zz < z; ++zz) out[zz * bs] = (x && pr < 0 && c < n) ? x[(int64_t)zz * n + c] : 0.0;
        return;
    }
    if (identity && r > n) {                      // [ I | 0 ] block
        const double v = (c < n && c == r - n - 1) ? 1.0 : 0.0;
        for (int zz = 0; zz < z; ++zz) out[zz * bs] = v;
        return;
    }
    const int a = idx ? idx[pr] : pr, b = idx ? idx[pc] : pc;
    const double d = D[(int64_t)a * ldd + b];
    const bool diag = (pr == pc);
    for (int zz = 0; zz < z; ++zz) {
        double v = hp[zz] * drp_exp(d * hp[2 * z + zz]);
        if (diag) v += hp[z + zz];
        out[zz * bs] = v;
    }
    (void)n2;
}

// Same augmented layout from an explicit covariance batch K [z,n,n] (general MultivariateNormal).
__global__ void __launch_bounds__(256) cov_assemble_kernel(const double* __restrict__ K, int n, const double* __restrict__ x,
                                                           double* __restrict__ M, int64_t ld, int64_t bs)
{
    const int r = blockIdx.y, zz = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > r) return;
    double v;
    if (r < n) v = K[((int64_t)zz * n + r) * n + c];
    else if (r == n) v = c < n ? x[(int64_t)zz * n + c] : 0.0;
    else v = (c < n && c == r - n - 1) ? 1.0 : 0.0;
    M[zz * bs + (int64_t)r * ld + c] = v;
}

// logp[z] = -1/2 (n log 2pi + y^T y) - sum_i log L_ii   (torch MultivariateNormal.log_prob)
__global__ void __launch_bounds__(256) mvn_finish_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                         double* __restrict__ logp)
{
    __shared__ double red[32];
    const double* A = M + (int64_t)blockIdx.x * bs;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += log(A[(int64_t)i * ld + i]);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        const double maha = -A[(int64_t)n * ld + n];
        logp[blockIdx.x] = -0.5 * ((double)n * LOG_2PI + maha) - s;
    }
}

// alpha[z][i] = -M[n+1+i][n]
__global__ void alpha_extract_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n, int cnt,
                                     double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out[(int64_t)blockIdx.y * cnt + i] = -M[(int64_t)blockIdx.y * bs + (int64_t)(n + 1 + i) * ld + n];
}

// ---------------------------------------------------------------------------------------------------------------
// Second phase of the two-phase prior: the factorisation (L, -K^-1: everything that costs time) depends on the
// hyper-parameters only and can start before the latent vector x exists; once x is there, alpha = K^-1 x is a symmetric
// matrix-vector product with the lower-stored -K^-1 block.  One CTA per 64 x 64 tile (I >= J): y_I += T x_J and, off the
// diagonal, y_J += T^T x_I.  Every (row block, other block) pair is produced exactly once -> ypart [z][nb][n], summed in
// a fixed order afterwards (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SYMV_TB = 64;
__global__ void __launch_bounds__(128) ou_symv_tile_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                           const double* __restrict__ x, int nb, double* __restrict__ ypart)
{
    __shared__ double T[SYMV_TB][SYMV_TB + 1];
    __shared__ double xi[SYMV_TB], xj[SYMV_TB];
    const int zz = blockIdx.y;
    int I = 0, t = blockIdx.x;                     // tile index -> (I, J), J <= I
    while (t > I) {
        t -= I + 1;
        ++I;
    }
    const int J = t;
    const double* A = M + (int64_t)zz * bs + (int64_t)(n + 1 + I * SYMV_TB) * ld + n + 1 + J * SYMV_TB;
    const int ri = min(SYMV_TB, n - I * SYMV_TB), cj = min(SYMV_TB, n - J * SYMV_TB);
    for (int idx = threadIdx.x; idx < SYMV_TB * SYMV_TB; idx += blockDim.x) {
        const int r = idx / SYMV_TB, c = idx - r * SYMV_TB;
        const bool in = r < ri && c < cj && (I != J || c <= r);
        T[r][c] = in ? A[(int64_t)r * ld + c] : 0.0;
    }
    if (threadIdx.x < SYMV_TB) {
        const int k = threadIdx.x;
        xi[k] = k < ri ? x[(int64_t)zz * n + I * SYMV_TB + k] : 0.0;
        xj[k] = k < cj ? x[(int64_t)zz * n + J * SYMV_TB + k] : 0.0;
    }
    __syncthreads();
    const int k = threadIdx.x & (SYMV_TB - 1);
    if (threadIdx.x < SYMV_TB) {                   // rows of block I
        double s = 0.0;
        if (I == J) {
            for (int c = 0; c < SYMV_TB; ++c) s += (c <= k ? T[k][c] : T[c][k]) * xj[c];
        } else {
            for (int c = 0; c < SYMV_TB; ++c) s += T[k][c] * xj[c];
        }
        if (k < ri) ypart[((int64_t)zz * nb + J) * n + I * SYMV_TB + k] = s;
    } else if (I != J) {                           // rows of block J: transposed tile
        double s = 0.0;
        for (int r = 0; r < SYMV_TB; ++r) s += T[r][k] * xi[r];
        if (k < cj) ypart[((int64_t)zz * nb + I) * n + J * SYMV_TB + k] = s;
    }
}

// alpha_i = -sum_q ypart[q][i] (the block holds -K^-1); written to alpha[z][n] and, in the layout the gradient kernels read,
// to M[n+1+i][n] = -alpha_i and M[n][n] = -x^T alpha.
__global__ void __launch_bounds__(256) ou_symv_finish_kernel(double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                             const double* __restrict__ x, int nb, const double* __restrict__ ypart,
                                                             double* __restrict__ alpha)
{
    __shared__ double red[32];
    const int zz = blockIdx.x;
    double* A = M + (int64_t)zz * bs;
    double maha = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nb; ++q) s += ypart[((int64_t)zz * nb + q) * n + i];
        const double a = -s;
        if (alpha) alpha[(int64_t)zz * n + i] = a;
        A[(int64_t)(n + 1 + i) * ld + n] = -a;
        maha += a * x[(int64_t)zz * n + i];
    }
    maha = block_sum(maha, red);
    if (threadIdx.x == 0) A[(int64_t)n * ld + n] = -maha;
}

// K5: hyper-parameter gradient sums.  With G = 1/2 (alpha alpha^T - K^-1) = dlogp/dK:
//   part[z][r] = ( sum_c G_rc E_rc , sum_c G_rc E_rc D_rc , G_rr )   over the full row r (by symmetry 2x lower).
// K^-1 is read once from the factored matrix (-M[n+1+i][n+1+j], lower), E is recomputed from D.
// alpha comes from the dense [z][n] copy (coalesced; the column M[n+1+c][n] it was read from before costs a 32-byte sector
// per entry - three times the traffic of the K^-1 row itself).  One WARP per row, eight rows per CTA, the row walked with
// four independent loads in flight per lane: with one CTA per row (60,000 CTAs, <= 8 dependent load -> exp iterations per
// thread and a block reduction each) the kernel ran at 24 % of the DRAM rate (ncu, profiles/r02/ncu_last_c4.summary.txt);
// a CTA walking eight rows with all its threads one after the other was slower still.
// part [z][nblk][3], nblk = ceil(n / OU_GRAD_ROWS).
constexpr int OU_GRAD_ROWS = 8;
__global__ void __launch_bounds__(256) ou_grad_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                      const double* __restrict__ D, int64_t ldd,
                                                      const double* __restrict__ lam, const double* __restrict__ alpha,
                                                      double* __restrict__ part)
{
    __shared__ double red[3][8];
    const int zz = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * OU_GRAD_ROWS + w;
    const double* A = M + (int64_t)zz * bs;
    const double* al = alpha + (int64_t)zz * n;
    const double nil = -1.0 / lam[zz];
    double sE = 0.0, sED = 0.0, gd = 0.0;
    if (r < n) {
        const double ar = al[r];
        const double* krow = A + (int64_t)(n + 1 + r) * ld + n + 1;
        const double* drow = D + (int64_t)r * ldd;
#pragma unroll 4
        for (int c = lane; c <= r; c += 32) {
            const double g = 0.5 * (ar * al[c] + krow[c]);       // krow = -Kinv
            const double d = drow[c];
            const double wgt = (c == r) ? 1.0 : 2.0;
            const double t = wgt * g * drp_exp(d * nil);
            sE += t;
            sED += t * d;
            if (c == r) gd = g;
        }
    }
    sE = warp_sum(sE);
    sED = warp_sum(sED);
    gd = warp_sum(gd);
    if (lane == 0) red[0][w] = sE, red[1][w] = sED, red[2][w] = gd;
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[threadIdx.x][q];
        part[((int64_t)zz * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = v;
    }
}

// G[z][i][j] = gscale[z] * 1/2 (alpha_i alpha_j - Kinv_ij), full symmetric (gradient w.r.t. an explicit covariance).
__global__ void __launch_bounds__(256) mvn_grad_cov_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                           const double* __restrict__ gscale, double* __restrict__ G)
{
    const int r = blockIdx.y, zz = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double* A = M + (int64_t)zz * bs;
    const int hi = max(r, c), lo = min(r, c);
    const double ar = -A[(int64_t)(n + 1 + r) * ld + n], ac = -A[(int64_t)(n + 1 + c) * ld + n];
    const double mk = A[(int64_t)(n + 1 + hi) * ld + n + 1 + lo];
    G[((int64_t)zz * n + r) * n + c] = (gscale ? gscale[zz] : 1.0) * 0.5 * (ar * ac + mk);
}

// Symmetric extraction of the trailing block: out[z][i][j] = sign * M[n+1+max][n+1+min] + jitter.
__global__ void __launch_bounds__(256) block_extract_kernel(const double* __restrict__ M, int64_t ld, int64_t bs, int n,
                                                            int cnt, double sign, double jitter, double* __restrict__ out)
{
    const int r = blockIdx.y, zz = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cnt) return;
    const int hi = max(r, c), lo = min(r, c);
    out[((int64_t)zz * cnt + r) * cnt + c] =
        sign * M[(int64_t)zz * bs + (int64_t)(n + 1 + hi) * ld + n + 1 + lo] + jitter;
}

}  // namespace

// ------------------------------------------------- entry points -----------------------------------------------

DRP_API int64_t drp_ou_workspace_ld(int n, int n2) { return ((int64_t)n + 1 + n2 + 7) & ~(int64_t)7; }

// elements (doubles) of the augmented workspace for z matrices: rows n+1+n2, leading dimension as above
static inline int64_t ou_matrix_elems(int n, int n2, int z) { return (int64_t)z * (n + 1 + n2) * drp_ou_workspace_ld(n, n2); }

// The tail of the workspace (behind the z augmented matrices) holds the int8 slices of the tcgen05 trailing update.
DRP_API int64_t drp_ou_workspace_elems(int n, int n2, int z)
{
    return ou_matrix_elems(n, n2, z) + (drp_ozaki_workspace_bytes(n + 1 + n2, z) + 7) / 8;
}

DRP_API int drp_ou_cov_fwd_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam,
                               int z, double* K, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!D) return -1;
    if (n <= 0 || ldd < n) return -3;
    if (!sf || !sn || !lam) return -4;
    if (z <= 0) return -7;
    if (!K) return -8;
    dim3 g((n + 255) / 256, n);
    {
        DRP_LAUNCH(DRP_K_OU_COV, ((double)n * n * 8.0 * (z + 1)), st);
        ou_cov_fwd_kernel<<<g, 256, 3 * z * sizeof(double), (cudaStream_t)stream>>>(D, ldd, n, sf, sn, lam, z, K);
    }
    return drp_launch_status();
}

// out[z][3] = (sum G.E, sum G.E.D, trace G); the caller applies d_sf = 2 sf S0, d_lam = sf^2 S1 / lam^2, d_sn = 2 sn S2.
// part: workspace of z*n*3 doubles.
DRP_API int drp_ou_cov_bwd_f64(const double* D, int64_t ldd, int n, const double* lam, int z, const double* G, double* part,
                               double* out, void* stream)
{
    if (!D) return -1;
    if (n <= 0 || ldd < n) return -3;
    if (!lam) return -4;
    if (z <= 0) return -5;
    if (!G || !part || !out) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    {
        DRP_LAUNCH(DRP_K_OU_GRAD, ((double)n * n * 8.0 * (z + 1)), st);
        ou_cov_bwd_kernel<<<dim3(n, z), 256, 0, st>>>(D, ldd, n, lam, G, part);
    }
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 24.0), st);
        reduce3_kernel<<<z, 256, 0, st>>>(part, n, out);
    }
    return drp_launch_status();
}

// Fused OU prior: logp[z] = log N(x_z; 0, K_z(D; sf, sn, lam)) and, if want_grad, alpha = K^-1 x [z,n] and the
// gradient sums gsum[z][3] (see drp_ou_cov_bwd_f64) of dlogp/dK.  The covariance never exists as a separate tensor.
//   M    workspace of drp_ou_workspace_elems(n, want_grad ? n : 0, z) doubles
//   part workspace of z*n*3 doubles (want_grad only)
DRP_API int drp_ou_prior_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam,
                             const double* x, int z, int want_grad, double* logp, double* alpha, double* gsum, double* M,
                             double* part, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!D) return -1;
    if (n <= 0 || ldd < n) return -3;
    if (!sf || !sn || !lam || !x) return -4;
    if (z <= 0) return -8;
    if (!logp || !M) return -10;
    if (want_grad && (!alpha || !gsum || !part)) return -11;
    cudaStream_t st = (cudaStream_t)stream;
    const int n2 = want_grad ? n : 0;
    const int R = n + 1 + n2;
    const int64_t ld = drp_ou_workspace_ld(n, n2), bs = (int64_t)R * ld;
    dim3 g((R + 255) / 256, R);
    {
        DRP_LAUNCH(DRP_K_ASSEMBLE, ((double)z * R * (R + 1) * 4.0), st);
        ou_assemble_kernel<<<g, 256, 3 * z * sizeof(double), st>>>(D, ldd, nullptr, n, n2, 1, sf, sn, lam, x, z, M, ld, bs);
    }
    const DrpGate gate{nullptr, sf, sn, (double)n};
    int rc = drp_factor_run(M, ld, bs, z, n, R, want_grad ? n + 1 : 0, R, info, M + ou_matrix_elems(n, n2, z), drp_ozaki_workspace_bytes(R, z), gate, &aux, st);
    if (rc) return rc;
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 8.0), st);
        mvn_finish_kernel<<<z, 256, 0, st>>>(M, ld, bs, n, logp);
    }
    if (want_grad) {
        {
            DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 16.0), st);
            alpha_extract_kernel<<<dim3((n + 255) / 256, z), 256, 0, st>>>(M, ld, bs, n, n, alpha);
        }
        {
            DRP_LAUNCH(DRP_K_OU_GRAD, ((double)z * n * n * 4.0 + (double)n * n * 4.0), st);
            ou_grad_kernel<<<dim3((n + OU_GRAD_ROWS - 1) / OU_GRAD_ROWS, z), 256, 0, st>>>(M, ld, bs, n, D, ldd, lam, alpha, part);
        }
        {
            DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 24.0), st);
            reduce3_kernel<<<z, 256, 0, st>>>(part, (n + OU_GRAD_ROWS - 1) / OU_GRAD_ROWS, gsum);
        }
    }
    return drp_launch_status();
}

// Two-phase form of drp_ou_prior_f64 (want_grad = 1 layout).  Phase 1 needs the hyper-parameters only: assembly +
// factorisation, leaving L and -K^-1 in M.  Phase 2 needs the latent vectors: alpha = K^-1 x, log-densities and the
// hyper-parameter gradient sums.  Same results as the one-call form up to rounding (alpha through the explicit inverse).
DRP_API int drp_ou_prior_factor_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam,
                                    int z, double* M, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!D) return -1;
    if (n <= 0 || ldd < n) return -3;
    if (!sf || !sn || !lam) return -4;
    if (z <= 0) return -7;
    if (!M) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = n + 1 + n;
    const int64_t ld = drp_ou_workspace_ld(n, n), bs = (int64_t)R * ld;
    dim3 g((R + 255) / 256, R);
    {
        DRP_LAUNCH(DRP_K_ASSEMBLE, ((double)z * R * (R + 1) * 4.0), st);
        ou_assemble_kernel<<<g, 256, 3 * z * sizeof(double), st>>>(D, ldd, nullptr, n, n, 1, sf, sn, lam, nullptr, z, M, ld, bs);
    }
    const DrpGate gate{nullptr, sf, sn, (double)n};
    int rc = drp_factor_run(M, ld, bs, z, n, R, n + 1, R, info, M + ou_matrix_elems(n, n, z), drp_ozaki_workspace_bytes(R, z), gate, &aux, st);
    if (rc) return rc;
    return drp_launch_status();
}

// Doubles of scratch (ypart) drp_ou_prior_apply_f64 needs.
DRP_API int64_t drp_ou_prior_apply_workspace_elems(int n, int z)
{
    const int64_t nb = (n + SYMV_TB - 1) / SYMV_TB;
    return (int64_t)z * nb * n;
}

DRP_API int drp_ou_prior_apply_f64(const double* D, int64_t ldd, int n, const double* lam, const double* x, int z, double* M,
                                   double* logp, double* alpha, double* gsum, double* part, double* ypart, void* stream)
{
    if (!D) return -1;
    if (n <= 0 || ldd < n) return -3;
    if (!lam || !x) return -4;
    if (z <= 0) return -6;
    if (!M) return -7;
    if (!logp || !alpha || !gsum || !part || !ypart) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = n + 1 + n;
    const int64_t ld = drp_ou_workspace_ld(n, n), bs = (int64_t)R * ld;
    const int nb = (n + SYMV_TB - 1) / SYMV_TB;
    {
        DRP_LAUNCH(DRP_K_SYMV, ((double)z * n * n * 4.0), st);
        ou_symv_tile_kernel<<<dim3(nb * (nb + 1) / 2, z), 128, 0, st>>>(M, ld, bs, n, x, nb, ypart);
    }
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * nb * n * 8.0), st);
        ou_symv_finish_kernel<<<z, 256, 0, st>>>(M, ld, bs, n, x, nb, ypart, alpha);
    }
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 8.0), st);
        mvn_finish_kernel<<<z, 256, 0, st>>>(M, ld, bs, n, logp);
    }
    {
        DRP_LAUNCH(DRP_K_OU_GRAD, ((double)z * n * n * 4.0 + (double)n * n * 4.0), st);
        ou_grad_kernel<<<dim3((n + OU_GRAD_ROWS - 1) / OU_GRAD_ROWS, z), 256, 0, st>>>(M, ld, bs, n, D, ldd, lam, alpha, part);
    }
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 24.0), st);
        reduce3_kernel<<<z, 256, 0, st>>>(part, (n + OU_GRAD_ROWS - 1) / OU_GRAD_ROWS, gsum);
    }
    return drp_launch_status();
}

// General batched MVN log-prob from an explicit covariance K [z,n,n] (zero mean handled by the caller: pass x - mu).
// want_grad: alpha [z,n] and Kinv-based G are available afterwards through drp_mvn_grad_cov_f64 on the same M.
// cond_bound (nullable, device [z]): upper bounds of cond_2(K_z); matrices within drp_ozaki_cond_limit take the
// split-integer tensor-core update, all others (and all of them without a bound) the fp64 one.
DRP_API int drp_mvn_logprob_f64(const double* K, const double* x, int n, int z, int want_grad, double* logp, double* alpha,
                                double* M, int32_t* info, const double* cond_bound, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!K || !x) return -1;
    if (n <= 0) return -3;
    if (z <= 0) return -4;
    if (!logp || !M) return -6;
    if (want_grad && !alpha) return -7;
    cudaStream_t st = (cudaStream_t)stream;
    const int n2 = want_grad ? n : 0;
    const int R = n + 1 + n2;
    const int64_t ld = drp_ou_workspace_ld(n, n2), bs = (int64_t)R * ld;
    {
        DRP_LAUNCH(DRP_K_ASSEMBLE, ((double)z * R * (R + 1) * 4.0 + (double)z * n * n * 4.0), st);
        cov_assemble_kernel<<<dim3((R + 255) / 256, R, z), 256, 0, st>>>(K, n, x, M, ld, bs);
    }
    const DrpGate gate{cond_bound, nullptr, nullptr, 0.0};
    int rc = drp_factor_run(M, ld, bs, z, n, R, want_grad ? n + 1 : 0, R, info, M + ou_matrix_elems(n, n2, z), drp_ozaki_workspace_bytes(R, z), gate, &aux, st);
    if (rc) return rc;
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 8.0), st);
        mvn_finish_kernel<<<z, 256, 0, st>>>(M, ld, bs, n, logp);
    }
    if (want_grad) {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 16.0), st);
        alpha_extract_kernel<<<dim3((n + 255) / 256, z), 256, 0, st>>>(M, ld, bs, n, n, alpha);
    }
    return drp_launch_status();
}

// G[z,n,n] = gscale[z] * dlogp_z/dK_z from a workspace factored by drp_mvn_logprob_f64 / drp_ou_prior_f64 (want_grad=1).
DRP_API int drp_mvn_grad_cov_f64(const double* M, int n, int z, const double* gscale, double* G, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!M) return -1;
    if (n <= 0 || z <= 0) return -2;
    if (!G) return -5;
    const int R = 2 * n + 1;
    const int64_t ld = drp_ou_workspace_ld(n, n), bs = (int64_t)R * ld;
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * n * 12.0), st);
        mvn_grad_cov_kernel<<<dim3((n + 255) / 256, n, z), 256, 0, (cudaStream_t)stream>>>(M, ld, bs, n, gscale, G);
    }
    return drp_launch_status();
}

// K^-1 of z SPD matrices via the same elimination (potri equivalent); Kinv [z,n,n] full symmetric.
DRP_API int drp_potri_batched_f64(const double* K, int n, int z, double* Kinv, double* M, int32_t* info,
                                  const double* cond_bound, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!K) return -1;
    if (n <= 0 || z <= 0) return -2;
    if (!Kinv || !M) return -4;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = 2 * n + 1;
    const int64_t ld = drp_ou_workspace_ld(n, n), bs = (int64_t)R * ld;
    cudaMemsetAsync(Kinv, 0, sizeof(double) * (size_t)z * n, st);   // borrowed as the zero right-hand side x
    {
        DRP_LAUNCH(DRP_K_ASSEMBLE, ((double)z * R * (R + 1) * 4.0 + (double)z * n * n * 4.0), st);
        cov_assemble_kernel<<<dim3((R + 255) / 256, R, z), 256, 0, st>>>(K, n, Kinv, M, ld, bs);
    }
    const DrpGate gate{cond_bound, nullptr, nullptr, 0.0};
    int rc = drp_factor_run(M, ld, bs, z, n, R, n + 1, R, info, M + ou_matrix_elems(n, n, z), drp_ozaki_workspace_bytes(R, z), gate, &aux, st);
    if (rc) return rc;
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * n * 12.0), st);
        block_extract_kernel<<<dim3((n + 255) / 256, n, z), 256, 0, st>>>(M, ld, bs, n, n, -1.0, 0.0, Kinv);
    }
    return drp_launch_status();
}

// Conditional posterior of the OU process at the `ni` nodes int_idx given values x_leaf [z,n] at the nodes leaf_idx:
//   mean [z,ni] = K_ab K_bb^-1 x_b,   cov [z,ni,ni] = K_aa - K_ab K_bb^-1 K_ba + jitter   (cov nullable: mean only)
// Mathematically identical to the precision-partition form of S/models.py:167-193 (Bishop B.49-50).
//   idx = [leaf_idx (n) | int_idx (ni)] rows of D;  M workspace of drp_ou_workspace_elems(n, ni, z) doubles.
DRP_API int drp_ou_conditional_f64(const double* D, int64_t ldd, const int32_t* idx, int n, int ni, const double* sf,
                                   const double* sn, const double* lam, const double* x_leaf, int z, double jitter,
                                   double* mean, double* cov, double* M, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream)
{
    const DrpAux aux{(cudaStream_t)aux_stream, (cudaEvent_t)ev_fork, (cudaEvent_t)ev_join};
    if (!D) return -1;
    if (!idx) return -3;
    if (n <= 0 || ni <= 0) return -4;
    if (!sf || !sn || !lam || !x_leaf) return -6;
    if (z <= 0) return -10;
    if (!mean || !M) return -12;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = n + 1 + ni;
    const int64_t ld = drp_ou_workspace_ld(n, ni), bs = (int64_t)R * ld;
    {
        DRP_LAUNCH(DRP_K_ASSEMBLE, ((double)z * R * (R + 1) * 4.0), st);
        ou_assemble_kernel<<<dim3((R + 255) / 256, R), 256, 3 * z * sizeof(double), st>>>(D, ldd, idx, n, ni, 0, sf, sn, lam,
                                                                                  x_leaf, z, M, ld, bs);
    }
    const DrpGate gate{nullptr, sf, sn, (double)(n + ni)};
    int rc = drp_factor_run(M, ld, bs, z, n, R, 0, cov ? R : n + 1, info, M + ou_matrix_elems(n, ni, z), drp_ozaki_workspace_bytes(R, z), gate, &aux, st);
    if (rc) return rc;
    {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * 16.0), st);
        alpha_extract_kernel<<<dim3((ni + 255) / 256, z), 256, 0, st>>>(M, ld, bs, n, ni, mean);
    }
    if (cov) {
        DRP_LAUNCH(DRP_K_REDUCE, ((double)z * n * n * 12.0), st);
        block_extract_kernel<<<dim3((ni + 255) / 256, ni, z), 256, 0, st>>>(M, ld, bs, n, ni, 1.0, jitter, cov);
    }
    return drp_launch_status();
}
