// K10: per-site categorical log-likelihood over the alignment, fused forward and backward.
//
// Replaces  pyro.sample("aa_sequences", dist.Categorical(logits=logits), obs=aminoacid_sequences)
// (S/models.py:352, S/ = /root/reference/draupnir/src/draupnir/) together with the decoder's LogSoftmax
// (S/models_utils.py:98):  sum_{i,l} ( s[i,l,x_il] - logsumexp_a s[i,l,a] ).  Since log_softmax is idempotent the
// kernel can be fed either the raw scores of `linear_probs` or already-normalised logits.
//
// HBM-bound: rows of A (=21) doubles are staged through shared memory with fully coalesced flat loads, one thread
// then owns one site.  The observations arrive as fp64 (the reference keeps residues in the fp64 dataset tensor and
// casts with .long() inside log_prob) or as int64.
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"

namespace {

constexpr int ROWS_PER_CTA = 256;

template <typename ObsT>
__global__ void __launch_bounds__(256) cat_fwd_kernel(const double* __restrict__ s, const ObsT* __restrict__ obs, int64_t rows,
                                                      int A, double* __restrict__ lse_out, double* __restrict__ part)
{
    extern __shared__ double tile[];   // [ROWS_PER_CTA][A] (+1 pad per row when A is even)
    __shared__ double red[32];
    const int lda = A | 1;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
    const int nrow = (int)min((int64_t)ROWS_PER_CTA, rows - r0);
    const double* src = s + r0 * A;
    // flat, fully coalesced copy of the CTA's rows; eight loads per thread are issued before the first store so that the
    // copy is bound by bandwidth instead of by the latency of one load per iteration
    const int total = nrow * A;
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * 256) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = i0 + u * 256;
            v[u] = idx < total ? src[idx] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = i0 + u * 256;
            if (idx < total) {
                const int r = idx / A, a = idx - r * A;
                tile[r * lda + a] = v[u];
            }
        }
    }
    __syncthreads();
    double ll = 0.0;
    if ((int)threadIdx.x < nrow) {
        const double* row = tile + threadIdx.x * lda;
        double m = row[0];
        for (int a = 1; a < A; ++a) m = fmax(m, row[a]);
        double e = 0.0;
        for (int a = 0; a < A; ++a) e += drp_exp(row[a] - m);
        const double lse = m + log(e);
        // an observation outside [0, A) (a code the alphabet does not have, a negative or NaN entry) poisons the sum with
        // NaN instead of reading a neighbouring row: the reference's gather would trip a device assert there
        const int x = (int)obs[r0 + threadIdx.x];
        ll = (x >= 0 && x < A) ? row[x] - lse : __longlong_as_double(0x7ff8000000000000LL);
        if (lse_out) lse_out[r0 + threadIdx.x] = lse;
    }
    ll = block_sum(ll, red);
    if (threadIdx.x == 0) part[blockIdx.x] = ll;
}

__global__ void __launch_bounds__(256) sum_kernel(const double* __restrict__ part, int n, double* __restrict__ out)
{
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// d/ds = g * (onehot(x) - softmax(s)), g = upstream gradient of the summed log-likelihood (device scalar)
template <typename ObsT>
__global__ void __launch_bounds__(256) cat_bwd_kernel(const double* __restrict__ s, const ObsT* __restrict__ obs,
                                                      const double* __restrict__ lse, const double* __restrict__ g,
                                                      int64_t rows, int A, double* __restrict__ ds)
{
    const int64_t total = rows * A;
    const double gv = g[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        double sv[4], lv[4];
        int ov[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // four independent element loads in flight per thread
            const int64_t idx = i0 + u * stride;
            if (idx < total) {
                const int64_t r = idx / A;
                av[u] = (int)(idx - r * A);
                sv[u] = s[idx];
                lv[u] = lse[r];
                ov[u] = (int)obs[r];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = i0 + u * stride;
            if (idx < total) {
                const double d = gv * ((ov[u] == av[u] ? 1.0 : 0.0) - drp_exp(sv[u] - lv[u]));
                ds[idx] = (ov[u] >= 0 && ov[u] < A) ? d : __longlong_as_double(0x7ff8000000000000LL);
            }
        }
    }
}

}  // namespace

DRP_API int64_t drp_cat_loglik_partials(int64_t rows) { return (rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA; }

// scores [rows, A] fp64; obs [rows] fp64 (obs_is_f64) or int64; out: 1 double; lse [rows] (nullable);
// part: workspace of drp_cat_loglik_partials(rows) doubles.
DRP_API int drp_cat_loglik_fwd_f64(const double* scores, const void* obs, int obs_is_f64, int64_t rows, int A, double* out,
                                   double* lse, double* part, void* stream)
{
    if (!scores || !obs) return -1;
    if (rows <= 0) return -4;
    if (A <= 0 || A > 64) return -5;
    if (!out || !part) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (int)drp_cat_loglik_partials(rows);
    const size_t sm = (size_t)ROWS_PER_CTA * (A | 1) * sizeof(double);
    if (drp_first_on_device(56)) {
        cudaFuncSetAttribute(cat_fwd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 65 * 8);
        cudaFuncSetAttribute(cat_fwd_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 65 * 8);
    }
    if (obs_is_f64) {
        DRP_LAUNCH(DRP_K_CATLIK, ((double)rows * (A * 8.0 + 16.0)), st);
        cat_fwd_kernel<double><<<nb, 256, sm, st>>>(scores, (const double*)obs, rows, A, lse, part);
    }
    else cat_fwd_kernel<int64_t><<<nb, 256, sm, st>>>(scores, (const int64_t*)obs, rows, A, lse, part);
    {
        DRP_LAUNCH(DRP_K_CATLIK, ((double)nb * 8.0), st);
        sum_kernel<<<1, 256, 0, st>>>(part, nb, out);
    }
    return drp_launch_status();
}

DRP_API int drp_cat_loglik_bwd_f64(const double* scores, const void* obs, int obs_is_f64, const double* lse, const double* g,
                                   int64_t rows, int A, double* dscores, void* stream)
{
    if (!scores || !obs || !lse || !g) return -1;
    if (rows <= 0) return -6;
    if (A <= 0) return -7;
    if (!dscores) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = rows * A;
    int64_t nb64 = (total + 255) / 256;
    const int nb = nb64 > 148 * 16 ? 148 * 16 : (int)nb64;
    if (obs_is_f64) {
        DRP_LAUNCH(DRP_K_CATLIK, ((double)rows * (A * 16.0 + 16.0)), st);
        cat_bwd_kernel<double><<<nb, 256, 0, st>>>(scores, (const double*)obs, lse, g, rows, A, dscores);
    }
    else cat_bwd_kernel<int64_t><<<nb, 256, 0, st>>>(scores, (const int64_t*)obs, lse, g, rows, A, dscores);
    return drp_launch_status();
}
