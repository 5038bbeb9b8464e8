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
int drp_factor_run(double* M, int64_t ld, int64_t bs, int z, int nfactor, int R_full, int grow_base, int col_limit,
                   int32_t* info, cudaStream_t st);
