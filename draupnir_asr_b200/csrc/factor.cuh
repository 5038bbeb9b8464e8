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
//   gate       which matrices may take the split-integer update (see DrpGate)
//
// Precision gate.  The split-integer update has a backward error of ~eps_S |K| (eps_6 ~ 2e-11, measured), fp64 LAPACK
// ~1e-16 |K|.  What the log-density and its gradients see is eps * cond(K), so a matrix only takes the tensor path
// when an UPPER BOUND of its condition number is below drp_ozaki_cond_limit() (1e-5 / eps_S; DRP_OZAKI_COND_LIMIT
// overrides); the others run every trailing update on the fp64 mma.sync kernel, in the same launches.  The decision is
// taken per matrix ON THE DEVICE (no host read of the hyper-parameters): a one-thread kernel writes the two index lists
// into the workspace and the update kernels read them.
struct DrpGate {
    const double* cond;      // [z] device: upper bounds of cond_2(K_z); or null
    const double* sf;        // else: OU hyper-parameters [z] (device) and the number of nodes `nnorm`, giving the bound
    const double* sn;        //       (nnorm * sf^2 + sn^2) / sn^2   (lambda_min(K) >= sn^2, |K|_2 <= trace(K))
    double nnorm;            // all null: no bound known -> fp64 updates only
};
// Look-ahead resources, owned by the caller (the library creates no streams or events): an auxiliary stream and two events.
struct DrpAux {
    cudaStream_t stream;
    cudaEvent_t ev_fork, ev_join;
};
int drp_factor_run(double* M, int64_t ld, int64_t bs, int z, int nfactor, int R_full, int grow_base, int col_limit,
                   int32_t* info, void* ws, int64_t ws_bytes, const DrpGate& gate, const DrpAux* aux, cudaStream_t st);

// ozaki.cu
int64_t drp_ozaki_workspace_bytes(int R_full, int z);
int drp_ozaki_slices();
double drp_ozaki_cond_limit();
// slice-buffer geometry of one update's panel (ozaki.cu: drp_ozaki_layout)
struct DrpOzLayout {
    uint8_t* sl;        // [z][S][KBN][Rpad][128 B], rows in the 128-byte-swizzled K-major image tcgen05 reads
    int64_t sl_bs;      // bytes per matrix
    int32_t* scl;       // [z][Rpad] binary exponent of every row's maximum
    int* counter;       // two work-unit counters, zeroed by whoever makes the slices
    int rb, Rpad, KBN;  // first sliced row (c1 rounded down to 8), rows (multiple of 128), 128-column k-blocks
};
bool drp_ozaki_layout(int z, int K, int c1, int R, int col_hi, int S, void* ws, int64_t ws_bytes, int buffer, DrpOzLayout* out);
// zmap (device, nullable = all z matrices): [0] number of matrices on the tensor path, [1] number on the fp64 path,
// [2 .. 2+z) indices of the former, [2+z .. 2+2z) indices of the latter
int64_t drp_zmap_bytes(int z);
int64_t drp_linv_bytes(int z);
int drp_ozaki_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_lo, int col_hi, int S,
                   void* ws, int64_t ws_bytes, const int* zmap, int do_slice, int buffer, int counter_idx, cudaStream_t st);
// factor.cu: the fp64 mma.sync trailing update (also exported for kernel-vs-kernel parity tests); with a zmap only the
// matrices of its fp64 list are updated
void drp_dmma_syrk(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit, const int* zmap,
                   cudaStream_t st);
// cudaFuncSetAttribute is per device: true the first time `key` is seen on the current device
bool drp_first_on_device(int key);
