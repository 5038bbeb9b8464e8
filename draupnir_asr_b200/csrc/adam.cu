// ClippedAdam update of ALL parameters of an SVI step in two launches, with the step counters and the learning rate on the
// device — the part of `svi.step` that follows the backward pass (S/main.py:1036-1041: pyro.optim.ClippedAdam with
// clip_norm 10, lrd 1; S/train.py:75).  pyro-ppl's ClippedAdam per parameter and step:
//     lr   <- lr * lrd                                   (once per optimiser call)
//     g    <- clamp(grad, -clip, +clip)  [+ weight_decay * p]          (element-wise clamp, despite the name)
//     m    <- b1 m + (1 - b1) g ;  v <- b2 v + (1 - b2) g^2
//     p    <- p - lr sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)
// Keeping t and lr on the device makes the update a fixed sequence of launches with fixed arguments, i.e. capturable in
// a CUDA graph together with the rest of the step and free of the ~10 multi-tensor launches per step of the eager form.
//
// The parameter list travels BY VALUE in the kernel arguments (a <= 4 KB __grid_constant__ struct): no device-side table
// that would have to be re-uploaded when the gradient buffers move (they do, every step, outside a graph).
#include "common.cuh"
#include "instrument.cuh"

namespace {

constexpr int ADAM_MAX = 72;          // tensors per launch: 72 * 48 B + header < 4 KB of kernel parameters

struct AdamSlot {
    double* p;
    const double* g;
    double* m;
    double* v;
    int64_t n;
    int64_t idx;                      // row of the state table: state[1 + 2 idx] = step count, state[2 + 2 idx] = step size
};

struct AdamBatch {
    AdamSlot s[ADAM_MAX];
    int count;
    int first;                        // first batch of this optimiser call: applies lr <- lr * lrd
};

// state[0] = lr.  One block; thread t handles tensor t of the batch.
__global__ void __launch_bounds__(128) adam_tick_kernel(const __grid_constant__ AdamBatch b, double* __restrict__ state, double lrd,
                                                        double b1, double b2)
{
    const int t = threadIdx.x;
    const double lr = state[0] * (b.first ? lrd : 1.0);
    __syncthreads();                                  // everyone has read the old lr
    if (t == 0 && b.first) state[0] = lr;
    if (t < b.count) {
        double* st = state + 1 + 2 * b.s[t].idx;
        const double step = st[0] + 1.0;
        st[0] = step;
        st[1] = lr * sqrt(1.0 - pow(b2, step)) / (1.0 - pow(b1, step));
    }
}

__global__ void __launch_bounds__(256) adam_update_kernel(const __grid_constant__ AdamBatch b, const double* __restrict__ state,
                                                          double b1, double b2, double eps, double wd, double clip)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
    for (int t = 0; t < b.count; ++t) {
        const AdamSlot s = b.s[t];
        const double step_size = state[2 + 2 * s.idx];
        for (int64_t i = gtid; i < s.n; i += gsz) {
            double g = s.g[i];
            g = g < -clip ? -clip : (g > clip ? clip : g);          // NaN passes through, as torch.clamp does
            const double p = s.p[i];
            if (wd != 0.0) g = fma(wd, p, g);
            const double m = s.m[i] * b1 + (1.0 - b1) * g;
            const double v = s.v[i] * b2 + (1.0 - b2) * g * g;
            s.m[i] = m;
            s.v[i] = v;
            s.p[i] = p - step_size * (m / (sqrt(v) + eps));
        }
    }
}

}  // namespace

// Largest number of tensors one call takes (the caller splits longer lists; only the first call of an optimiser step
// passes first = 1).
DRP_API int drp_clipped_adam_max_tensors(void) { return ADAM_MAX; }

// ptrs: HOST array of 4 * count device pointers (p, grad, exp_avg, exp_avg_sq per tensor); numel / slot: HOST arrays of
// count entries (elements; row of `state`).  state: DEVICE doubles [1 + 2 * slots]: lr, then (step count, step size) per
// slot — initialise lr and zero the rest.  Enqueues two kernels on `stream`; nothing is read back.
DRP_API int drp_clipped_adam_f64(const void* const* ptrs, const int64_t* numel, const int64_t* slot, int count, int first,
                                 double* state, double lrd, double beta1, double beta2, double eps, double weight_decay,
                                 double clip, void* stream)
{
    if (!ptrs || !numel || !slot) return -1;
    if (count <= 0 || count > ADAM_MAX) return -4;
    if (!state) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    AdamBatch b;
    int64_t total = 0;
    for (int t = 0; t < count; ++t) {
        b.s[t].p = (double*)ptrs[4 * t];
        b.s[t].g = (const double*)ptrs[4 * t + 1];
        b.s[t].m = (double*)ptrs[4 * t + 2];
        b.s[t].v = (double*)ptrs[4 * t + 3];
        if (!b.s[t].p || !b.s[t].g || !b.s[t].m || !b.s[t].v) return -1;
        if (numel[t] < 0 || slot[t] < 0) return -2;
        b.s[t].n = numel[t];
        b.s[t].idx = slot[t];
        total += numel[t];
    }
    b.count = count;
    b.first = first ? 1 : 0;
    {
        DRP_LAUNCH(DRP_K_OTHER, 32.0 * count, st);
        adam_tick_kernel<<<1, 128, 0, st>>>(b, state, lrd, beta1, beta2);
    }
    {
        int64_t blocks = (total + 1023) / 1024;                // ~4 elements per thread
        if (blocks < 1) blocks = 1;
        if (blocks > 592) blocks = 592;                        // 4 CTAs per SM
        DRP_LAUNCH(DRP_K_OTHER, 56.0 * (double)total, st);
        adam_update_kernel<<<(int)blocks, 256, 0, st>>>(b, state, beta1, beta2, eps, weight_decay, clip);
    }
    return drp_launch_status();
}
