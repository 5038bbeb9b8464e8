// K1: all-pairs patristic (path-length) and cladistic (edge-count) distances + LCA indices on the GPU.
//
// Replaces the offline builder of the reference (S/ = /root/reference/draupnir/src/draupnir/):
//   calculate_patristic_distance            S/utils.py:496-579   (O(N^2) ete3 get_distance loop, n <= 200)
//   Rscript Calculate_Patristic.R           S/Calculate_Patristic.R:9   (ape::dist.nodes, n > 200)
//
// Nodes are BFS level-order ids (root 0, parent[v] < v; S/utils.py:864).  d(i,j) = r_i + r_j - 2 r_lca(i,j) with
// r = distance to the root.  The LCA of (i, j) is the first node on i's ancestor chain whose subtree (a contiguous
// preorder interval) contains j, found by binary search over the chain held in shared memory: O(log depth) per pair,
// no N x log N tables, output written once, coalesced, in the reference's row-major level-order layout.
#include "common.cuh"
#include "instrument.cuh"
#include "factor.cuh"

namespace {

// Tree preparation: depth, root distance, subtree size and preorder index of every node (every subtree becomes the
// contiguous preorder interval [pre, pre + size)).  Sequential reference path: one lane, three passes; parent[v] < v
// makes one forward pass enough.  Used for orderings that are topological but not level-order, and for N > 32768.
__device__ void tree_prepare_sequential(const int32_t* __restrict__ parent, const double* __restrict__ branch, int N,
                                        int32_t* __restrict__ depth, double* __restrict__ root_dist,
                                        int32_t* __restrict__ pre, int32_t* __restrict__ size, int32_t* __restrict__ next_free)
{
    depth[0] = 0;
    root_dist[0] = 0.0;
    for (int v = 1; v < N; ++v) {
        const int p = parent[v];
        depth[v] = depth[p] + 1;
        root_dist[v] = root_dist[p] + branch[v];
    }
    for (int v = 0; v < N; ++v) size[v] = 1;
    for (int v = N - 1; v >= 1; --v) size[parent[v]] += size[v];
    pre[0] = 0;
    next_free[0] = 1;
    for (int v = 1; v < N; ++v) {
        const int p = parent[v];
        pre[v] = next_free[p];
        next_free[p] += size[v];
        next_free[v] = pre[v] + 1;
    }
}

constexpr int TP_THREADS = 1024;
constexpr int TP_PER_THREAD = 32;       // the parallel path covers N <= 32768 nodes

// Parallel path for the reference's numbering (BFS level order, S/utils.py:864: nodes sorted by depth, siblings
// adjacent): depths by pointer jumping (log2(depth) rounds), then one bottom-up sweep (subtree sizes) and one top-down
// sweep (root distances - same summation order as the sequential path, so bit-identical - and preorder indices) over
// the levels, each level processed by the whole CTA.  Any other numbering falls back to the sequential path.
__global__ void __launch_bounds__(TP_THREADS) tree_prepare_kernel(const int32_t* __restrict__ parent,
                                                                  const double* __restrict__ branch, int N,
                                                                  int32_t* __restrict__ depth, double* __restrict__ root_dist,
                                                                  int32_t* __restrict__ pre, int32_t* __restrict__ size,
                                                                  int32_t* __restrict__ scratch)
{
    __shared__ int s_any, s_bad;
    const int tid = threadIdx.x;
    if (N > TP_THREADS * TP_PER_THREAD) {
        if (tid == 0) tree_prepare_sequential(parent, branch, N, depth, root_dist, pre, size, scratch);
        return;
    }
    // ---- depths: anc (in `pre`) and the edge count to it (in `depth`) double every round
    for (int v = tid; v < N; v += TP_THREADS) {
        pre[v] = v ? parent[v] : -1;
        depth[v] = v ? 1 : 0;
        scratch[v] = 0;
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int round = 0;; ++round) {
        if (tid == 0) s_any = 0;
        __syncthreads();
        int na[TP_PER_THREAD], nc[TP_PER_THREAD];
        int any = 0;
#pragma unroll
        for (int k = 0; k < TP_PER_THREAD; ++k) {
            const int v = tid + k * TP_THREADS;
            na[k] = -2;
            if (v < N) {
                const int a = pre[v];
                if (a >= 0) {
                    na[k] = pre[a];
                    nc[k] = depth[a];
                    any = 1;
                }
            }
        }
        __syncthreads();                                   // every read of this round precedes every write
#pragma unroll
        for (int k = 0; k < TP_PER_THREAD; ++k) {
            const int v = tid + k * TP_THREADS;
            if (na[k] != -2) {
                pre[v] = na[k];
                depth[v] += nc[k];
            }
        }
        if (any) s_any = 1;
        __syncthreads();
        const int more = s_any;
        __syncthreads();                                   // everyone has read the flag before it is reset
        if (!more) break;
        if (round > 40) {                                  // 2^40 levels cannot exist: parent[] is not a forest
            if (tid == 0) s_bad = 1;
            break;
        }
    }
    __syncthreads();
    // ---- is this a level order with adjacent siblings?
    for (int v = tid + 1; v < N; v += TP_THREADS) {
        if (depth[v] < depth[v - 1]) s_bad = 1;
        if (parent[v] != parent[v - 1]) atomicAdd(&scratch[parent[v]], 1);      // one block of children per parent
    }
    __syncthreads();
    for (int v = tid; v < N; v += TP_THREADS)
        if (scratch[v] > 1) s_bad = 1;
    __syncthreads();
    if (s_bad) {
        if (tid == 0) tree_prepare_sequential(parent, branch, N, depth, root_dist, pre, size, scratch);
        return;
    }
    // ---- level boundaries: scratch[d] = first node of depth d
    const int maxd = depth[N - 1];
    for (int v = tid; v < N; v += TP_THREADS) {
        if (v == 0 || depth[v] != depth[v - 1]) scratch[depth[v]] = v;
        size[v] = 1;
    }
    __syncthreads();
    // ---- bottom-up: subtree sizes
    for (int d = maxd; d >= 1; --d) {
        const int lo = scratch[d], hi = d == maxd ? N : scratch[d + 1];
        for (int v = lo + tid; v < hi; v += TP_THREADS) atomicAdd(&size[parent[v]], size[v]);
        __syncthreads();
    }
    // ---- top-down: root distances and preorder indices
    if (tid == 0) {
        root_dist[0] = 0.0;
        pre[0] = 0;
    }
    __syncthreads();
    for (int d = 1; d <= maxd; ++d) {
        const int lo = scratch[d], hi = d == maxd ? N : scratch[d + 1];
        for (int v = lo + tid; v < hi; v += TP_THREADS) {
            const int p = parent[v];
            root_dist[v] = root_dist[p] + branch[v];
            if (parent[v - 1] != p) {                                       // first child lays out its siblings
                int run = pre[p] + 1;
                for (int w = v; w < N && parent[w] == p; ++w) {
                    pre[w] = run;
                    run += size[w];
                }
            }
        }
        __syncthreads();
    }
}

struct ChainEntry {
    int32_t lo, hi;    // preorder interval [lo, hi) of the ancestor's subtree
    int32_t id, depth;
    double rd;         // root distance of the ancestor
};

constexpr int CHAIN_SMEM = 1024;   // ancestors held in shared memory; deeper chains spill to the global workspace

// RPC requested row nodes per CTA.  The ancestors a_0 = i, a_1, ..., root of a row node are laid out in shared memory; the
// LCA of (i, j) is the deepest a_k whose subtree (the preorder interval [pre, pre + size)) contains j.
//   TABLE = true : a lookup table over PREORDER positions (uint16 chain index per position, 2 N bytes of shared memory
//                  per row) is filled once per row from the nested intervals - position p in subtree(a_k) \
//                  subtree(a_{k-1}) gets k - so a column costs two shared-memory reads instead of a binary search, and
//                  the per-column operands (pre[j], root_dist[j], depth[j]) are fetched once for the RPC rows of the CTA.
//   TABLE = false: one row per CTA, binary search over the chain per column (trees too large for the table, or chains
//                  deeper than 65535 / longer than the shared-memory chain).
template <bool TABLE, int RPC, int NT>
__global__ void __launch_bounds__(NT) lca_patristic_kernel(const int32_t* __restrict__ parent, const int32_t* __restrict__ depth,
                                                            const double* __restrict__ root_dist, const int32_t* __restrict__ pre,
                                                            const int32_t* __restrict__ size, const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ cols, int nr, int nc, int n_nodes,
                                                            int32_t* __restrict__ lca_out, double* __restrict__ dist_out,
                                                            double* __restrict__ clad_out, ChainEntry* __restrict__ spill,
                                                            int spill_stride)
{
    constexpr int CHAIN_PER_ROW = CHAIN_SMEM / RPC;
    __shared__ ChainEntry chain_s[CHAIN_SMEM];
    extern __shared__ uint16_t table[];            // [RPC][n_nodes] chain index by preorder position (TABLE only)
    int irow[RPC], pi[RPC], di0[RPC];
    double ri[RPC];
    ChainEntry* chain[RPC];
#pragma unroll
    for (int q = 0; q < RPC; ++q) {
        const int rix = min((int)blockIdx.x * RPC + q, nr - 1);            // tail CTA: repeat the last row (same values)
        irow[q] = rows ? rows[rix] : rix;
        pi[q] = pre[irow[q]];
        di0[q] = depth[irow[q]];
        ri[q] = root_dist[irow[q]];
        chain[q] = (TABLE || di0[q] < CHAIN_PER_ROW) ? chain_s + q * CHAIN_PER_ROW
                                                     : spill + (int64_t)(blockIdx.x * RPC + q) * spill_stride;
    }
    // ancestor k of i = the node of depth depth[i] - k whose preorder interval contains pre[i]: found by a parallel
    // containment test over all nodes (coalesced) instead of a dependent parent[] walk by one lane (len x DRAM latency)
    // (four nodes per thread and iteration, all loads issued before the first use: the kernel is bound by load latency)
    for (int a0 = threadIdx.x; a0 < n_nodes; a0 += 4 * NT) {
        int lo4[4], hi4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * NT;
            lo4[u] = a < n_nodes ? pre[a] : 0x7fffffff;
            hi4[u] = a < n_nodes ? size[a] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * NT;
            const int lo = lo4[u], hi = lo + hi4[u];
#pragma unroll
            for (int q = 0; q < RPC; ++q) {
                if (lo <= pi[q] && pi[q] < hi) {
                    ChainEntry e;
                    e.lo = lo;
                    e.hi = hi;
                    e.id = a;
                    e.depth = depth[a];
                    e.rd = root_dist[a];
                    chain[q][di0[q] - e.depth] = e;
                }
            }
        }
    }
    __syncthreads();
    if (!TABLE) __threadfence_block();
    if (TABLE) {
#pragma unroll
        for (int q = 0; q < RPC; ++q) {
            uint16_t* tb = table + (int64_t)q * n_nodes;
            int plo = 0, phi = 0;                  // interval of the previous (deeper) ancestor; empty for k = 0
            for (int k = 0; k <= di0[q]; ++k) {
                const int lo = chain[q][k].lo, hi = chain[q][k].hi;
                const int first = k == 0 ? hi : plo;   // k = 0: the whole interval [lo, hi)
                for (int p = lo + threadIdx.x; p < first; p += blockDim.x) tb[p] = (uint16_t)k;
                if (k > 0)
                    for (int p = phi + threadIdx.x; p < hi; p += blockDim.x) tb[p] = (uint16_t)k;
                plo = lo;
                phi = hi;
            }
        }
        __syncthreads();
    }
    for (int b0 = threadIdx.x; b0 < nc; b0 += 4 * NT) {
        int p4[4], dj4[4];
        double rj4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // all loads of four columns in flight before the first use
            const int b = b0 + u * NT;
            const int j = b < nc ? (cols ? cols[b] : b) : 0;
            p4[u] = pre[j];
            rj4[u] = root_dist[j];
            dj4[u] = clad_out ? depth[j] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * NT;
            if (b >= nc) break;
            const int p = p4[u];
#pragma unroll
            for (int q = 0; q < RPC; ++q) {
                const int rix = (int)blockIdx.x * RPC + q;
                if (rix >= nr) break;
                int k;
                if (TABLE) {
                    k = table[(int64_t)q * n_nodes + p];
                } else {
                    // smallest k with lo_k <= p < hi_k (intervals are nested, the root's always matches)
                    int lo = 0, hi = di0[q];
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const ChainEntry e = chain[q][mid];
                        if (e.lo <= p && p < e.hi) hi = mid;
                        else lo = mid + 1;
                    }
                    k = lo;
                }
                const ChainEntry e = chain[q][k];
                const int64_t o = (int64_t)rix * nc + b;
                if (lca_out) lca_out[o] = e.id;
                if (dist_out) dist_out[o] = (ri[q] - e.rd) + (rj4[u] - e.rd);
                if (clad_out) clad_out[o] = (double)(di0[q] + dj4[u] - 2 * e.depth);
            }
        }
    }
}

constexpr int TABLE_SMEM_MAX = 160 * 1024;         // dynamic shared memory for the preorder tables (2 bytes per node and row)

}  // namespace

// depth/root_dist/pre/size/scratch: device arrays of N elements each (scratch is clobbered).
// Requires only parent[v] < v (any topological numbering); the reference's level order takes the parallel path.
DRP_API int drp_tree_prepare(const int32_t* parent, const double* branch, int32_t n_nodes, int32_t* depth, double* root_dist,
                             int32_t* pre, int32_t* size, int32_t* scratch, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!parent || !branch) return -1;
    if (n_nodes <= 0) return -3;
    if (!depth || !root_dist || !pre || !size || !scratch) return -4;
    {
        DRP_LAUNCH(DRP_K_PATRISTIC, ((double)n_nodes * 40.0), st);
        tree_prepare_kernel<<<1, TP_THREADS, 0, (cudaStream_t)stream>>>(parent, branch, n_nodes, depth, root_dist, pre, size, scratch);
    }
    return drp_launch_status();
}

// bytes of spill workspace needed for `nr` rows when the deepest node has `max_depth` edges to the root
DRP_API int64_t drp_lca_workspace_bytes(int32_t nr, int32_t max_depth)
{
    return max_depth + 1 <= CHAIN_SMEM ? 0 : (int64_t)nr * (max_depth + 1) * (int64_t)sizeof(ChainEntry);
}

// rows/cols: node ids to pair up (nullable = all nodes 0..nr-1 / 0..nc-1).  Any of the three outputs may be null.
DRP_API int drp_lca_patristic_f64(const int32_t* parent, const int32_t* depth, const double* root_dist, const int32_t* pre,
                                  const int32_t* size, int32_t n_nodes, int32_t max_depth, const int32_t* rows, int32_t nr,
                                  const int32_t* cols, int32_t nc, int32_t* lca_out, double* dist_out, double* clad_out,
                                  void* ws, int64_t ws_bytes, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!parent || !depth || !root_dist || !pre || !size) return -1;
    if (n_nodes <= 0) return -6;
    if (nr <= 0 || nc <= 0) return -9;
    if (ws_bytes < drp_lca_workspace_bytes(nr, max_depth) || (ws_bytes > 0 && !ws)) return -16;
    {
        DRP_LAUNCH(DRP_K_PATRISTIC, ((double)nr * nc * ((dist_out ? 8.0 : 0.0) + (clad_out ? 8.0 : 0.0) + (lca_out ? 4.0 : 0.0))), st);
        if (drp_first_on_device(58)) {
            cudaFuncSetAttribute(lca_patristic_kernel<true, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_SMEM_MAX);
            cudaFuncSetAttribute(lca_patristic_kernel<true, 1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_SMEM_MAX);
        }
        const size_t row_bytes = (size_t)n_nodes * sizeof(uint16_t);
        // two rows per CTA of 512 threads: the per-column operands and the ancestor scan (both L2 traffic, which is what
        // bounds this kernel) are shared by the two rows at the same number of resident threads per SM
        if (max_depth < CHAIN_SMEM / 2 && 2 * row_bytes <= 80 * 1024 && nr >= 2) {
            lca_patristic_kernel<true, 2, 512><<<(nr + 1) / 2, 512, 2 * row_bytes, st>>>(parent, depth, root_dist, pre, size, rows,
                                                                                      cols, nr, nc, n_nodes, lca_out, dist_out,
                                                                                      clad_out, (ChainEntry*)ws, max_depth + 1);
        } else if (max_depth < CHAIN_SMEM && row_bytes <= (size_t)TABLE_SMEM_MAX) {
            lca_patristic_kernel<true, 1, 256><<<nr, 256, row_bytes, st>>>(parent, depth, root_dist, pre, size, rows, cols, nr, nc,
                                                                           n_nodes, lca_out, dist_out, clad_out, (ChainEntry*)ws,
                                                                           max_depth + 1);
        } else {
            lca_patristic_kernel<false, 1, 256><<<nr, 256, 0, st>>>(parent, depth, root_dist, pre, size, rows, cols, nr, nc,
                                                                    n_nodes, lca_out, dist_out, clad_out, (ChainEntry*)ws,
                                                                    max_depth + 1);
        }
    }
    return drp_launch_status();
}
