// Launch accounting + optional per-kernel-class CUDA-event timing (used by bench.py for `gpu_launches` and the live
// roofline).  Disabled timing costs one relaxed atomic add per launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum DrpKernelClass {
    DRP_K_SYRK = 0,      // trailing update (fp64 tensor pipe)          work = flops
    DRP_K_TRSM = 1,      // panel solve                                 work = flops
    DRP_K_POTF2 = 2,     // diagonal block                              work = flops
    DRP_K_ASSEMBLE = 3,  // OU covariance / augmented matrix assembly   work = bytes
    DRP_K_REDUCE = 4,    // log-prob / gradient reductions / extraction work = bytes
    DRP_K_CATLIK = 5,    // categorical log-likelihood fwd/bwd          work = bytes
    DRP_K_PATRISTIC = 6, // LCA / patristic                             work = bytes
    DRP_K_OTHER = 7,
    DRP_K_SYRK_TC = 8,   // trailing update on the int8 tcgen05 pipe         work = fp64-equivalent flops
    DRP_K_SLICE = 9,     // fp64 panel -> int8 slices for the tcgen05 update work = bytes
    DRP_K_GRU = 10,      // bidirectional GRU recurrence fwd / bwd                work = recurrent-GEMM flops
    DRP_K_CLASSES = 11
};

// Call right before / after a kernel launch of class `k` doing `work` algorithmic flops or bytes.
void drp_instr_begin(int k, double work, cudaStream_t st);
void drp_instr_end(int k, cudaStream_t st);

struct DrpScopedLaunch {
    int k;
    cudaStream_t st;
    DrpScopedLaunch(int k_, double work, cudaStream_t st_) : k(k_), st(st_) { drp_instr_begin(k, work, st); }
    ~DrpScopedLaunch() { drp_instr_end(k, st); }
};
#define DRP_CAT2(a, b) a##b
#define DRP_CAT(a, b) DRP_CAT2(a, b)
#define DRP_LAUNCH(klass, work, st) DrpScopedLaunch DRP_CAT(_drp_scope_, __LINE__)((klass), (double)(work), (st))
