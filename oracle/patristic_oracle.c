/* TEST INFRASTRUCTURE (see oracle/__init__.py) — scalar CPU restatement of the patristic builder.
 *
 * Follows the reference's ete3 branch of calculate_patristic_distance
 * (/root/reference/draupnir/src/draupnir/utils.py:572-575): for every pair (t1, t2)
 * tree.get_distance(t1, t2) walks from each node up to the common ancestor and adds the branch lengths
 * along both arms (topology_only=True counts the edges instead -> cladistic matrix).  ete3 itself is not
 * vendored under /root/reference and is absent from this image, so its published behaviour is restated:
 * the distance is accumulated arm by arm, NOT as r_i + r_j - 2 r_lca, which makes this an independent
 * check of the product's formula (they agree to rounding, rtol 1e-12).
 *
 * Nodes are BFS level-order ids (root 0), so parent[v] < v and the deeper node always has the larger
 * depth; the walk lifts the deeper node until both meet.  The full matrix is returned (the reference fills
 * the upper triangle and symmetrises later with np.maximum, utils.py:1183-1186).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>

void drp_oracle_patristic(const int32_t *parent, const double *branch, const int32_t *depth, int32_t n_nodes,
                          const int32_t *rows, int32_t nr, const int32_t *cols, int32_t nc,
                          int32_t *lca_out, double *patristic_out, double *cladistic_out)
{
    (void)n_nodes;
    for (int32_t a = 0; a < nr; ++a) {
        for (int32_t b = 0; b < nc; ++b) {
            int32_t u = rows[a], v = cols[b];
            double du = 0.0, dv = 0.0;
            int32_t edges = 0;
            while (u != v) {
                if (depth[u] >= depth[v]) { du += branch[u]; u = parent[u]; }
                else                      { dv += branch[v]; v = parent[v]; }
                ++edges;
            }
            int64_t o = (int64_t)a * nc + b;
            if (lca_out) lca_out[o] = u;
            if (patristic_out) patristic_out[o] = du + dv;
            if (cladistic_out) cladistic_out[o] = (double)edges;
        }
    }
}
