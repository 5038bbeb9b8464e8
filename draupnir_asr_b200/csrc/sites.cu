// Summed log-densities of the small sample sites of the model / guide, one launch each way.
//
// Replaces, per site and ELBO evaluation, the 15-30 element-wise torch kernels (and as many again in the backward pass) behind
//   pyro.sample("alpha" | "sigma_f" | "sigma_n" | "lambd", dist.HalfNormal(scale).expand_by([k]).to_event(1))   S/models.py:89-92
//   pyro.sample("latent_z", dist.Normal(z_loc.T, z_scale.T))  (the guide's q(z))                               S/guides.py:118-121
// Below ~800 leaves and on 8 GPUs those ~300 launches of a few microseconds each are a large share of the step (a 19-leaf
// step: 270 of 306 kernels), captured CUDA graph or not.
//   HalfNormal(s).log_prob(v) = log 2 - 1/2 log(2 pi) - log s - v^2 / (2 s^2)      (v >= 0; -inf otherwise)
//   Normal(m, s).log_prob(v)  =        - 1/2 log(2 pi) - log s - (v - m)^2 / (2 s^2)
// Sums are deterministic: per-block partials in a fixed order, then one block adds them.
#include "common.cuh"
#include "instrument.cuh"

namespace {

constexpr double HALF_LOG_2PI = 0.91893853320467274178032973640562;
constexpr double LOG_2 = 0.69314718055994530941723212145818;
constexpr int SITE_THREADS = 256;
constexpr int SITE_MAX_BLOCKS = 256;

// Operands are logical [rows, cols] arrays read through (row stride, column stride) in elements: 0 broadcasts, (1, rows) is a
// transposed view (the guide hands over z_loc.T / z_scale.T), so no operand is ever copied to be made contiguous.
struct Strided {
    const double* p;
    int64_t sr, sc;
};
__device__ __forceinline__ double at(const Strided& a, int64_t r, int64_t c) { return a.p[r * a.sr + c * a.sc]; }

// mode 0: HalfNormal (loc ignored), 1: Normal.
__global__ void __launch_bounds__(SITE_THREADS) site_logp_kernel(Strided v, Strided m, Strided s, int64_t n, int64_t cols, int mode,
                                                                 double* __restrict__ part)
{
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const double x = at(v, r, c), sc = at(s, r, c);
        const double d = mode ? x - at(m, r, c) : x;
        const double q = d / sc;
        double lp = -HALF_LOG_2PI - log(sc) - 0.5 * q * q;
        if (!mode) lp = x >= 0.0 ? lp + LOG_2 : -INFINITY;
        acc += lp;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(SITE_THREADS) site_sum_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out)
{
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = acc;
}

// Gradients of g * sum(log_prob): dv, dm (Normal only, nullable), ds — dense [n] each (a broadcast operand's gradient is
// reduced by the caller's autograd, as for any expanded tensor).
__global__ void __launch_bounds__(SITE_THREADS) site_logp_bwd_kernel(Strided v, Strided m, Strided s, int64_t n, int64_t cols, int mode,
                                                                     const double* __restrict__ g, double* __restrict__ dv,
                                                                     double* __restrict__ dm, double* __restrict__ ds)
{
    const double gv = g[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const double x = at(v, r, c), sc = at(s, r, c);
        const double d = mode ? x - at(m, r, c) : x;
        const double is = 1.0 / sc;
        const double q = d * is;
        const double t = -q * is * gv;                 // d lp / d v
        if (dv) dv[i] = t;
        if (dm) dm[i] = -t;
        if (ds) ds[i] = (q * q - 1.0) * is * gv;       // d lp / d s = -1/s + d^2 / s^3
    }
}

}  // namespace

// Doubles of workspace (`part`) the forward call needs.
DRP_API int64_t drp_site_logp_workspace_elems(void) { return SITE_MAX_BLOCKS; }

// out[0] = sum over a logical [rows, cols] array of log_prob(value).  mode 0: HalfNormal(scale), 1: Normal(loc, scale).
// strides: HOST array of 6 element strides (value row, value col, loc row, loc col, scale row, scale col); 0 broadcasts.
DRP_API int drp_site_logp_f64(const double* value, const double* loc, const double* scale, const int64_t* strides, int64_t rows,
                              int64_t cols, int mode, double* out, double* part, void* stream)
{
    if (!value || !scale || (mode && !loc)) return -1;
    if (!strides) return -4;
    if (rows <= 0 || cols <= 0) return -5;
    if (mode != 0 && mode != 1) return -7;
    if (!out || !part) return -8;
    const int64_t n = rows * cols;
    const Strided v{value, strides[0], strides[1]}, m{loc, strides[2], strides[3]}, s{scale, strides[4], strides[5]};
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + SITE_THREADS - 1) / SITE_THREADS;
    if (blocks > SITE_MAX_BLOCKS) blocks = SITE_MAX_BLOCKS;
    {
        DRP_LAUNCH(DRP_K_REDUCE, (double)n * 24.0, st);
        site_logp_kernel<<<(int)blocks, SITE_THREADS, 0, st>>>(v, m, s, n, cols, mode, blocks == 1 ? out : part);
    }
    if (blocks > 1) {
        DRP_LAUNCH(DRP_K_REDUCE, (double)blocks * 8.0, st);
        site_sum_kernel<<<1, SITE_THREADS, 0, st>>>(part, (int)blocks, out);
    }
    return drp_launch_status();
}

// g: device scalar (upstream gradient of the sum).  dvalue / dloc / dscale: dense row-major [rows, cols], each nullable.
DRP_API int drp_site_logp_bwd_f64(const double* value, const double* loc, const double* scale, const int64_t* strides,
                                  int64_t rows, int64_t cols, int mode, const double* g, double* dvalue, double* dloc,
                                  double* dscale, void* stream)
{
    if (!value || !scale || (mode && !loc)) return -1;
    if (!strides) return -4;
    if (rows <= 0 || cols <= 0) return -5;
    if (mode != 0 && mode != 1) return -7;
    if (!g) return -8;
    const int64_t n = rows * cols;
    const Strided v{value, strides[0], strides[1]}, m{loc, strides[2], strides[3]}, s{scale, strides[4], strides[5]};
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + SITE_THREADS - 1) / SITE_THREADS;
    if (blocks > 1184) blocks = 1184;
    DRP_LAUNCH(DRP_K_REDUCE, (double)n * 48.0, st);
    site_logp_bwd_kernel<<<(int)blocks, SITE_THREADS, 0, st>>>(v, m, s, n, cols, mode, g, dvalue, dloc, dscale);
    return drp_launch_status();
}
