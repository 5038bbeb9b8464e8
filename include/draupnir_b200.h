/* draupnir_b200 — C ABI of libdraupnir_b200.so (sm_100a).
 *
 * The drop-in boundary for DRAUPNIR's per-step tree-OU prior + ELBO likelihood path.  The reference
 * (LysSanzMoreta/DRAUPNIR_ASR) is pure Python and has no FFI of its own; each entry point below replaces the torch
 * call(s) cited next to it (S/ = draupnir/src/draupnir/ in the reference tree) and is what a ctypes stub in the
 * reference would bind (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into row-major contiguous storage unless a leading dimension is given;
 *   - the caller owns every buffer including workspaces (sizes from the *_workspace_* helpers); the library
 *     allocates nothing and keeps no state that affects results (the only process-wide data are the launch counter /
 *     opt-in CUDA-event timing of the instrumentation block below and the "opt-in shared memory configured on this
 *     device" bits), so calls are re-entrant per stream and CUDA-graph capturable;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 ok, -k = argument group k invalid, 1000+e = CUDA runtime error e;
 *   - `info[z]` (int32, device): 0, or the 1-based index of the first non-positive pivot of matrix z (LAPACK potrf
 *     convention).  The reference's behaviour on that condition is torch.linalg.LinAlgError; the Python wrapper
 *     raises the same class.
 *   - fp64 ("_f64") is the reference dtype (Draupnir_example.py:146,154).
 */
#ifndef DRAUPNIR_B200_H
#define DRAUPNIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version (bumped on any signature change). */
int drp_abi_version(void);

/* ---- instrumentation (bench.py: `gpu_launches` and the live per-kernel roofline) -------------------------------- */
/* Cumulative number of kernels launched by this library.                                                            */
int64_t drp_launch_count(void);
/* CUDA-event timing of every launch, by kernel class (one class per kernel of the path; drp_timing_class_name(k) names it,
 * drp_timing_class_is_flops(k) says whether work[k] counts flops or bytes).  enable(on >= 1) clears earlier records
 * (on > 1 additionally pre-creates the events for `on` launches, keeping event creation out of the caller's timed
 * region); read() synchronises the device and fills three arrays of drp_timing_classes() entries (HOST pointers).
 * Not to be enabled while a CUDA graph is being captured.  This block is the library's only process-wide state.       */
int drp_timing_classes(void);
const char* drp_timing_class_name(int k);   /* the kernel behind class k ("oz_syrk_kernel", "gru_fwd2_kernel", ...)    */
int drp_timing_class_is_flops(int k);       /* 1: work[k] counts flops, 0: bytes                                        */
int drp_timing_enable(int on);
int drp_timing_read(double* ms, int64_t* launches, double* work);

/* ---- K1  patristic builder: S/utils.py:496-579 (ete3 get_distance loop) and S/Calculate_Patristic.R:9 --------- */
/* Tree given as parent[] / branch[] with parent[v] < v (level-order ids, root 0, S/utils.py:864).                  */
int drp_tree_prepare(const int32_t* parent, const double* branch, int32_t n_nodes, int32_t* depth, double* root_dist,
                     int32_t* pre, int32_t* size, int32_t* scratch, void* stream);
int64_t drp_lca_workspace_bytes(int32_t nr, int32_t max_depth);
/* rows/cols nullable (= all nodes).  lca_out int32 [nr,nc], dist_out / clad_out fp64 [nr,nc], each nullable.       */
int drp_lca_patristic_f64(const int32_t* parent, const int32_t* depth, const double* root_dist, const int32_t* pre,
                          const int32_t* size, int32_t n_nodes, int32_t max_depth, const int32_t* rows, int32_t nr,
                          const int32_t* cols, int32_t nc, int32_t* lca_out, double* dist_out, double* clad_out, void* ws,
                          int64_t ws_bytes, void* stream);

/* ---- K2  OUKernel_Fast.forward: S/models_utils.py:303-311 ------------------------------------------------------ */
/* D [n,n] with leading dimension ldd (so patristic_matrix_sorted[1:,1:] needs no copy); K [z,n,n].                 */
int drp_ou_cov_fwd_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam, int z,
                       double* K, void* stream);
/* out[z][3] = (sum G.E, sum G.E.D, trace G) for G = dLoss/dK; part: z*n*3 doubles of workspace.                    */
int drp_ou_cov_bwd_f64(const double* D, int64_t ldd, int n, const double* lam, int z, const double* G, double* part,
                       double* out, void* stream);

/* ---- K3  torch.linalg.cholesky inside MultivariateNormal.__init__: S/models.py:118 ----------------------------- */
/* ws (nullable): drp_factor_workspace_bytes(n, z) bytes; with it, trailing updates of large matrices run on the int8
 * tcgen05 tensor pipe (split-integer; DRP_OZAKI_SLICES = 0 disables, 4..7 selects the slice count) — but only for the
 * matrices whose condition number is known to be small enough for its ~2e-11 |K| backward error to stay invisible:
 * cond_bound (nullable, DEVICE [z]) holds upper bounds of cond_2(K_z); matrix z takes the tensor path iff
 * cond_bound[z] <= drp_ozaki_cond_limit() (1e-5 / 2^-(8S-13): 3.4e5 for S = 6; DRP_OZAKI_COND_LIMIT overrides), decided
 * on the device.  Every other matrix — and every matrix when cond_bound is NULL — gets fp64 (mma.sync) updates, i.e. the
 * accuracy of the LAPACK routine the reference calls.  The OU entry points below derive the bound from the
 * hyper-parameters themselves: (nodes * sf^2 + sn^2) / sn^2.
 * With a workspace every 128-column panel is also factored by two fused kernels (whole diagonal region in one CTA per
 * matrix; row-independent blocked substitution of the rows below on the fp64 tensor pipe, with one refinement step for the
 * matrices off the tensor path) instead of a chain of 64-column potf2 / trsm / update steps; the workspace's tail holds the
 * inverses of the 64 x 64 diagonal blocks (DRP_PANEL_WIDTH = 0 keeps the chain).                                       */
int64_t drp_factor_workspace_bytes(int rows, int z);
double drp_ozaki_cond_limit_f64(void);
/* Look-ahead (every entry point that factorises takes the same three arguments): aux_stream, ev_fork, ev_join — a second
 * cudaStream_t and two cudaEvent_t of the caller's (the library creates none), or NULL.  With them the update that leaves
 * an outer panel is split by columns: the next panel's columns first, on `stream`; the rest on aux_stream, concurrently with
 * the next panel's latency-bound factorisation chain.  aux_stream is joined back into `stream` before the call returns
 * (stream-wise), so the caller sees one stream; the construct is CUDA-graph capturable.  DRP_LOOKAHEAD=0 disables.  */
int drp_potrf_batched_f64(double* A, int n, int z, int zero_upper, int32_t* info, const double* cond_bound, void* ws,
                          int64_t ws_bytes, void* aux_stream, void* ev_fork, void* ev_join, void* stream);
/* Test hook: ONE trailing update C[r][c] -= sum_{k0 <= k < k0+K} P[r][k] P[c][k], rows [c1,R), cols [c1,min(R,col_limit)),
 * c <= r; slices = 0 -> fp64 mma.sync kernel, 4..7 -> int8 tcgen05 kernel.  Returns 1 if tcgen05 rejects the shape.  */
int drp_syrk_update_f64(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_limit, int slices,
                        void* ws, int64_t ws_bytes, void* stream);
/* Test hook: only the columns [col_lo, col_hi) of that update (what the look-ahead splits an update into); do_slice = 0
 * re-uses the slices made by the preceding call for the same panel, counter_idx (0/1) picks the work-unit counter.     */
int drp_syrk_update_range_f64(double* M, int64_t ld, int64_t bs, int z, int k0, int K, int c1, int R, int col_lo, int col_hi,
                              int slices, void* ws, int64_t ws_bytes, int do_slice, int counter_idx, void* stream);

/* ---- K3+K4+K5  gp_prior's latent_z site, fused: S/models.py:100-118 -------------------------------------------- */
int64_t drp_ou_workspace_ld(int n, int n2);
int64_t drp_ou_workspace_elems(int n, int n2, int z);
/* logp[z] = log N(x_z; 0, sf_z^2 exp(-D/lam_z) + sn_z^2 I).  want_grad: alpha[z,n] = K^-1 x and gsum[z][3]
 * (as drp_ou_cov_bwd_f64 with G = dlogp/dK).  M: drp_ou_workspace_elems(n, want_grad ? n : 0, z) doubles.          */
int drp_ou_prior_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam,
                     const double* x, int z, int want_grad, double* logp, double* alpha, double* gsum, double* M,
                     double* part, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream);

/* The same in two phases (want_grad = 1 layout, M of drp_ou_workspace_elems(n, n, z) doubles): the factorisation needs
 * only the hyper-parameters (S/models.py:90-92,101-105) and may be enqueued before the latent vectors exist (they come
 * out of the guide's encoder, S/guides.py:116-120); the second call needs x (S/models.py:118) and costs one symmetric
 * matrix-vector product per latent dimension.  ypart: drp_ou_prior_apply_workspace_elems(n, z) doubles.             */
int drp_ou_prior_factor_f64(const double* D, int64_t ldd, int n, const double* sf, const double* sn, const double* lam,
                            int z, double* M, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream);
int64_t drp_ou_prior_apply_workspace_elems(int n, int z);
int drp_ou_prior_apply_f64(const double* D, int64_t ldd, int n, const double* lam, const double* x, int z, double* M,
                           double* logp, double* alpha, double* gsum, double* part, double* ypart, void* stream);

/* ---- K4/K5  MultivariateNormal(0, K).log_prob(x) for an explicit covariance batch ------------------------------ */
/* cond_bound: as for drp_potrf_batched_f64 (NULL = fp64 updates only).                                               */
int drp_mvn_logprob_f64(const double* K, const double* x, int n, int z, int want_grad, double* logp, double* alpha,
                        double* M, int32_t* info, const double* cond_bound, void* aux_stream, void* ev_fork, void* ev_join,
                        void* stream);
int drp_mvn_grad_cov_f64(const double* M, int n, int z, const double* gscale, double* G, void* stream);
int drp_potri_batched_f64(const double* K, int n, int z, double* Kinv, double* M, int32_t* info, const double* cond_bound,
                          void* aux_stream, void* ev_fork, void* ev_join, void* stream);

/* ---- K6  conditional_sampling / conditional_samplingMAP: S/models.py:149-242 ----------------------------------- */
/* idx = [leaf rows of D (n) | internal rows of D (ni)]; mean [z,ni]; cov [z,ni,ni] nullable (+jitter everywhere).  */
int drp_ou_conditional_f64(const double* D, int64_t ldd, const int32_t* idx, int n, int ni, const double* sf,
                           const double* sn, const double* lam, const double* x_leaf, int z, double jitter, double* mean,
                           double* cov, double* M, int32_t* info, void* aux_stream, void* ev_fork, void* ev_join, void* stream);

/* ---- K9  nn.GRU (bidirectional, one layer, hidden 60) inside RNNDecoder_Tiling.forward / RNNEncoder.forward:
 *          S/models_utils.py:93-94 and :54 ----------------------------------------------------------------------- */
/* gi [n,L,2,3H] = W_ih x + b_ih per direction (gate order r,z,n); whh [2,3H,H]; bhh [2,3H]; h0 [2,n,H];
 * out [n,L,2H] (torch layout: forward direction first); gates [n,L,2,4,H] saved for the backward pass (nullable).
 * H must equal drp_gru_hidden_size() (60, the reference's gru_hidden_dim: S/main.py:2777).                        */
int drp_gru_hidden_size(void);
int drp_gru_fwd_f64(const double* gi, const double* whh, const double* bhh, const double* h0, int n, int L, int H,
                    double* out, double* gates, void* stream);
/* Forward direction only: gi0 [n,L,3H] compact; out / gates keep the two-direction layout and only their direction-0
 * halves are written (the encoder's loss path only needs output[:, -1]: S/models_utils.py:57).                    */
int drp_gru_fwd_dir0_f64(const double* gi0, const double* whh, const double* bhh, const double* h0, int n, int L, int H,
                         double* out, double* gates, void* stream);
/* Same with gi[i][l] = U[i] + V[l], U [n,2,3H], V [L,2,3H]: the decoder's tiled input [latent_i ; blosum_l]
 * (S/models.py:339-342) is never materialised.                                                                     */
int drp_gru_fwd_uv_f64(const double* U, const double* V, const double* whh, const double* bhh, const double* h0, int n,
                       int L, int H, double* out, double* gates, void* stream);
/* dgi [n,L,2,3H] = dLoss/dgi; dh0 [2,n,H] nullable.  dLoss/dW_hh = sum dgh^T h_prev with dgh = dgi, n-part times r.   */
int drp_gru_bwd_f64(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                    int n, int L, int H, double* dgi, double* dh0, void* stream);

/* Forward direction only: dgi0 [n,L,3H] compact, dh0[0 .. n*H) written.  For callers whose loss only sees the hidden
 * state at the last position (RNNEncoder.forward, S/models_utils.py:57): the reverse direction is then one cell step.  */
int drp_gru_bwd_dir0_f64(const double* dout, const double* out, const double* gates, const double* whh, const double* h0,
                         int n, int L, int H, double* dgi0, double* dh0, void* stream);
/* Same when the gradient enters at the last position only: dlast [n,2H] (direction-0 half read); spares the caller a
 * zero-filled [n,L,2H] tensor (RNNEncoder.forward, S/models_utils.py:57).                                             */
int drp_gru_bwd_last_dir0_f64(const double* dlast, const double* out, const double* gates, const double* whh,
                              const double* h0, int n, int L, int H, double* dgi0, double* dh0, void* stream);
/* Gradients of the decoder's broadcast input terms gi[i][l] = U[i] + V[l] (drp_gru_fwd_uv_f64) in one pass over
 * dgi [n,L,C]: dUpart [drp_gru_input_sums_chunks(n, L), n, C] = sums over the positions of a chunk, dVpart
 * [drp_gru_input_sums_blocks(n), L, C] = per-row-block sums over i; the caller adds the chunks / blocks up.  C even, C <= 512.                                                                                  */
int drp_gru_input_sums_blocks(int n);
int drp_gru_input_sums_chunks(int n, int L);
int drp_gru_input_sums_f64(const double* dgi, int n, int L, int C, double* dUpart, double* dVpart, void* stream);
int drp_gru_wgrad_dir0_f64(const double* dgi0, const double* gates, const double* out, const double* h0, int n, int L, int H,
                           double* dwhh, double* dbhh, double* ws, int64_t ws_elems, void* stream);
/* dwhh [2,3H,H] = sum_p dgh[p]^T h_prev[p], dbhh [2,3H] = sum_p dgh[p] (autograd of nn.GRU's weight_hh / bias_hh);
 * ws: drp_gru_wgrad_workspace_elems() doubles.                                                                   */
int64_t drp_gru_wgrad_workspace_elems(void);
int drp_gru_wgrad_f64(const double* dgi, const double* gates, const double* out, const double* h0, int n, int L, int H,
                      double* dwhh, double* dbhh, double* ws, int64_t ws_elems, void* stream);

/* ---- weight / bias gradients of the dense layers around the recurrences (autograd of nn.Linear: S/models_utils.py:54-65,
 *      93-99):  C [M,N] = A^T B, colsum [M] = sum_k A[k][:] (nullable) for A [K,M], B [K,N], M, N <= 192, K arbitrary ------ */
int64_t drp_gram_workspace_elems(int M, int N);
/* 1 if drp_gram_tn_f64 takes the shape (its 8x8 output tiles split into 8 warp blocks of <= 3 x 5, stages fit), else 0.   */
int drp_gram_tn_supported(int M, int N);
int drp_gram_tn_f64(const double* A, const double* B, int64_t K, int M, int N, double* C, double* colsum, double* ws,
                    int64_t ws_elems, void* stream);

/* ---- nn.Linear on n*L rows (EmbedComplex.fc1/fc2 S/models_utils.py:232-236, RNNDecoder_Tiling.fc1/linear_probs :97-98,
 *          the GRU input projection W_ih x + b_ih of RNNEncoder :54) and its input gradient --------------------------- */
/* Y [M,N] = X [M,K] Wt + bias: X row-major contiguous (16-byte aligned), Wt [K,N] = the TRANSPOSED weight, bias [N]
 * nullable.  M huge, K <= 192, N <= 184.  One read of X, one write of Y.  Returns 1 without launching when W^T and
 * the ring of X blocks exceed shared memory (K and N both large).                                                    */
int drp_tall_gemm_f64(const double* X, const double* Wt, const double* bias, int64_t M, int K, int N, double* Y,
                      void* stream);

/* ---- K10  Categorical(logits).log_prob(obs).sum(): S/models.py:352 (+ LogSoftmax S/models_utils.py:98) --------- */
int64_t drp_cat_loglik_partials(int64_t rows);
int drp_cat_loglik_fwd_f64(const double* scores, const void* obs, int obs_is_f64, int64_t rows, int A, double* out,
                           double* lse, double* part, void* stream);
int drp_cat_loglik_bwd_f64(const double* scores, const void* obs, int obs_is_f64, const double* lse, const double* g,
                           int64_t rows, int A, double* dscores, void* stream);

/* ---- the small sample sites, summed log-density in one launch each way:
 *      dist.HalfNormal(scale).expand_by([k]).to_event(1) at "alpha", "sigma_f", "sigma_n", "lambd" (S/models.py:89-92) and the
 *      guide's dist.Normal(z_loc.T, z_scale.T) at "latent_z" (S/guides.py:118-121).  mode 0: HalfNormal(scale) (loc unused),
 *      mode 1: Normal(loc, scale).  Operands are logical [rows, cols] arrays read through strides[6] (HOST array, elements:
 *      value row/col, loc row/col, scale row/col; 0 broadcasts, so expanded scalars and transposed views need no copy).
 *      part: drp_site_logp_workspace_elems() doubles.  bwd: g = device scalar; dvalue / dloc / dscale dense [rows, cols],
 *      each nullable.                                                                                                    */
int64_t drp_site_logp_workspace_elems(void);
int drp_site_logp_f64(const double* value, const double* loc, const double* scale, const int64_t* strides, int64_t rows,
                      int64_t cols, int mode, double* out, double* part, void* stream);
int drp_site_logp_bwd_f64(const double* value, const double* loc, const double* scale, const int64_t* strides, int64_t rows,
                          int64_t cols, int mode, const double* g, double* dvalue, double* dloc, double* dscale, void* stream);

/* ---- pyro.optim.ClippedAdam over all parameters of the step: S/main.py:1036-1041 (the update that ends svi.step,
 *      S/train.py:75).  lr <- lr * lrd; g = clamp(grad, +-clip) + weight_decay * p; Adam moments; p -= lr sqrt(1 - b2^t) /
 *      (1 - b1^t) * m / (sqrt(v) + eps).  Step counters and lr live on the DEVICE (state [1 + 2 * slots] doubles: lr, then
 *      (t, step size) per slot; the caller sets state[0] = lr and zeroes the rest), so the update is a fixed launch
 *      sequence: CUDA-graph capturable.  ptrs: HOST array of 4 * count device pointers (p, grad, exp_avg, exp_avg_sq per
 *      tensor); numel, slot: HOST arrays [count].  count <= drp_clipped_adam_max_tensors(); longer lists are split by the
 *      caller, first = 1 only for the first part (it applies lrd).                                                      */
int drp_clipped_adam_max_tensors(void);
int drp_clipped_adam_f64(const void* const* ptrs, const int64_t* numel, const int64_t* slot, int count, int first,
                         double* state, double lrd, double beta1, double beta2, double eps, double weight_decay, double clip,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRAUPNIR_B200_H */
