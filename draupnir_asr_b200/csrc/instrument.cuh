// Launch accounting + optional per-kernel-class CUDA-event timing (used by bench.py for `gpu_launches` and the live
// roofline).  Disabled timing costs one relaxed atomic add per launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum DrpKernelClass {
    DRP_K_SYRK = 0,      // syrk_kernel / syrk_gated_kernel: trailing update, fp64 mma.sync          work = flops
    DRP_K_TRSM = 1,      // trsm_kernel: panel solve                                              work = flops
    DRP_K_POTF2 = 2,     // potf2_kernel: diagonal block                                          work = flops
    DRP_K_ASSEMBLE = 3,  // ou_assemble_kernel / cov_assemble_kernel: augmented matrix assembly        work = bytes
    DRP_K_REDUCE = 4,    // small reductions / extractions (mvn_finish, alpha_extract, reduce3 ...) work = bytes
    DRP_K_CATLIK = 5,    // cat_fwd_kernel / cat_bwd_kernel: categorical log-likelihood           work = bytes
    DRP_K_PATRISTIC = 6, // tree_prepare_kernel + lca_patristic_kernel                            work = bytes
    DRP_K_OTHER = 7,     // gate_kernel, zero_upper_kernel, adam_*_kernel                          work = bytes
    DRP_K_SYRK_TC = 8,   // oz_syrk_kernel: trailing update on the int8 tcgen05 pipe              work = fp64-equivalent flops
    DRP_K_SLICE = 9,     // oz_slice_kernel: fp64 panel -> int8 slices                            work = bytes
    DRP_K_GRU_FWD = 10,  // gru_fwd2_kernel (gru_fwd_kernel): recurrence forward                  work = recurrent-GEMM flops
    DRP_K_GRU_BWD = 11,  // gru_bwd_kernel: back-propagation through time                         work = recurrent-GEMM flops
    DRP_K_GRU_WGRAD = 12,// gru_wgrad_kernel: weight_hh / bias_hh gradients                       work = flops
    DRP_K_TALL_GEMM = 13,// tall_gemm_kernel: nn.Linear on n*L rows                               work = bytes
    DRP_K_GRAM = 14,     // gram_tn_kernel (+ gram_reduce_kernel): tall-skinny weight gradients   work = bytes
    DRP_K_GRU_SUMS = 15, // gru_input_sums_kernel: dU / dV of the decoder's broadcast input       work = bytes
    DRP_K_OU_GRAD = 16,  // ou_grad_kernel / ou_cov_bwd_kernel: hyper-parameter gradient sums     work = bytes
    DRP_K_SYMV = 17,     // ou_symv_tile_kernel: alpha = K^-1 x from the stored inverse           work = bytes
    DRP_K_OU_COV = 18,   // ou_cov_fwd_kernel: dense [z,n,n] OU covariance (OUKernel_Fast.forward)            work = bytes
    DRP_K_POTRF_DIAG = 19,  // potrf_diag_kernel: W x W diagonal region of a panel (potf2 + in-region solves/updates)   work = flops
    DRP_K_PANEL_SOLVE = 20, // panel_solve_kernel: blocked forward substitution of the rows below, fp64 mma.sync        work = flops
    DRP_K_CLASSES = 21
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
