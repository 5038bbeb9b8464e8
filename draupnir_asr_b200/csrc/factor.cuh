// Internal interface of the partial-Cholesky engine (factor.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DRP_NB 64

// Eliminate the first `nfactor` columns of z lower-stored matrices (row-major, `ld`, batch stride `bs`).
//   R_full     rows/cols of the whole augmented matrix
//   grow_base  0: every row < R_full is active from the first panel;
//              >0: rows >= grow_base + (columns eliminated so far) are still exactly zero in the panel and are
//                  skipped (identity right-hand-side block of the OU prior)
//   col_limit  trailing updates only touch columns < col_limit
//   info       [z] int32, 0 or 1-based index of the first non-positive pivot (LAPACK convention); may be null
//   ws         optional workspace (drp_ozaki_workspace_bytes) enabling the int8 tcgen05 trailing update; without it
//              (or for small trailing matrices) the update runs on the fp64 mma.sync kernel
int drp_factor_run(double* M, int64_t ld, int64_t bs, int z, int nfactor, int R_full, int grow_base, int col_limit,
                   int32_t* info, void* ws, int64_t ws_bytes, cudaStream_t st);

// ozaki.cu
int64_t drp_ozaki_workspace_bytes(int R_full, int z);
int drp_ozaki_slices();
int drp_ozaki_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit, int S, void* ws,
                   int64_t ws_bytes, cudaStream_t st);
// factor.cu: the fp64 mma.sync trailing update (also exported for kernel-vs-kernel parity tests)
void drp_dmma_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit, cudaStream_t st);
