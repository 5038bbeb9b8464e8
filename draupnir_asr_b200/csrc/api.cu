// ABI bookkeeping + instrumentation for libdraupnir_b200.so (see include/draupnir_b200.h).
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"
#include <atomic>
#include <mutex>
#include <vector>

namespace {
std::atomic<long long> g_launches[DRP_K_CLASSES];
std::atomic<int> g_timing{0};
std::mutex g_mu;
struct Span { cudaEvent_t a, b; };
std::vector<Span> g_spans[DRP_K_CLASSES];
double g_work[DRP_K_CLASSES];
cudaEvent_t g_pending[DRP_K_CLASSES];
std::vector<cudaEvent_t> g_free;      // recycled events: creating them inside a timed region stalls the driver now and then

cudaEvent_t take_event()
{
    cudaEvent_t e;
    if (!g_free.empty()) {
        e = g_free.back();
        g_free.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
}  // namespace

void drp_instr_begin(int k, double work, cudaStream_t st)
{
    g_launches[k].fetch_add(1, std::memory_order_relaxed);
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    g_work[k] += work;
    cudaEvent_t e = take_event();
    cudaEventRecord(e, st);
    g_pending[k] = e;
}

void drp_instr_end(int k, cudaStream_t st)
{
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t e = take_event();
    cudaEventRecord(e, st);
    g_spans[k].push_back({g_pending[k], e});
}

// cudaFuncSetAttribute (opt-in shared memory) is a per-device setting: the lazily applied attributes are tracked per
// (kernel key, device), so a process that drives several GPUs configures each of them.
bool drp_first_on_device(int key)
{
    static std::atomic<unsigned long long> seen[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (seen[key & 63].fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
}

DRP_API int drp_abi_version(void) { return 5; }

// Number of kernel classes reported by drp_timing_read (length of its three arrays).
DRP_API int drp_timing_classes(void) { return DRP_K_CLASSES; }

// Name of the (dominant) kernel of class k and whether its work is counted in flops (1) or bytes (0).
DRP_API const char* drp_timing_class_name(int k)
{
    static const char* names[DRP_K_CLASSES] = {
        "syrk_kernel", "trsm_kernel", "potf2_kernel", "ou_assemble_kernel", "small_reductions", "cat_fwd_bwd_kernel",
        "lca_patristic_kernel", "misc", "oz_syrk_kernel", "oz_slice_kernel", "gru_fwd2_kernel", "gru_bwd_kernel",
        "gru_wgrad_kernel", "tall_gemm_kernel", "gram_tn_kernel", "gru_input_sums_kernel", "ou_grad_kernel",
        "ou_symv_tile_kernel", "ou_cov_fwd_kernel", "potrf_diag_kernel", "panel_solve_kernel"};
    return (k >= 0 && k < DRP_K_CLASSES) ? names[k] : "";
}
DRP_API int drp_timing_class_is_flops(int k)
{
    return k == DRP_K_SYRK || k == DRP_K_TRSM || k == DRP_K_POTF2 || k == DRP_K_SYRK_TC || k == DRP_K_GRU_FWD ||
           k == DRP_K_GRU_BWD || k == DRP_K_GRU_WGRAD || k == DRP_K_POTRF_DIAG || k == DRP_K_PANEL_SOLVE;
}

// Cumulative number of kernels this library has launched (all classes).
DRP_API int64_t drp_launch_count(void)
{
    long long s = 0;
    for (int k = 0; k < DRP_K_CLASSES; ++k) s += g_launches[k].load();
    return (int64_t)s;
}

// Enable (on >= 1) / disable (0) CUDA-event timing of every launch; enabling clears previous records.  `on` > 1 also
// pre-creates (and pre-records) events for that many launches, so that the driver grows neither its event pool nor its
// timestamp pool inside the caller's timed region (either costs ~100 ms once every ~1000 events).
DRP_API int drp_timing_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int k = 0; k < DRP_K_CLASSES; ++k) {
        for (auto& s : g_spans[k]) { g_free.push_back(s.a); g_free.push_back(s.b); }
        g_spans[k].clear();
        g_work[k] = 0.0;
    }
    if (on > 1) {
        while ((long long)g_free.size() < 2ll * on) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) break;
            g_free.push_back(e);
        }
        // ... and records every event once: the driver attaches the timestamp storage of an event at its first record and
        // grows that pool in steps too (single 100-190 ms steps were seen when the 1024th / 2048th record of a run fell
        // into the timed region)
        cudaStream_t tmp;
        if (cudaStreamCreateWithFlags(&tmp, cudaStreamNonBlocking) == cudaSuccess) {
            for (cudaEvent_t e : g_free) cudaEventRecord(e, tmp);
            cudaStreamSynchronize(tmp);
            cudaStreamDestroy(tmp);
        }
    }
    g_timing.store(on ? 1 : 0);
    return 0;
}

// Synchronises the device and fills, per kernel class k < drp_timing_classes(): ms[k] summed device time,
// launches[k] timed launches, work[k] summed algorithmic flops (classes 0-2, 8) or bytes (classes 3-6, 9).
DRP_API int drp_timing_read(double* ms, int64_t* launches, double* work)
{
    if (!ms || !launches || !work) return -1;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_mu);
    for (int k = 0; k < DRP_K_CLASSES; ++k) {
        double t = 0.0;
        for (auto& s : g_spans[k]) {
            float f = 0.f;
            if (cudaEventElapsedTime(&f, s.a, s.b) == cudaSuccess) t += f;
        }
        ms[k] = t;
        launches[k] = (int64_t)g_spans[k].size();
        work[k] = g_work[k];
    }
    return 0;
}
