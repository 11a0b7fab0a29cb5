/* oracle/dp_haploid.c — TEST INFRASTRUCTURE ONLY (parity checker; never linked into the product).
 *
 * Plain-C restatement of the reference's haploid DP, Approximator::dp_approximation_solver
 * (src/approximator.cpp:44-168), in the reference's own evaluation order:
 *   push relaxation over the Kahn-ordered expanded graph (:55-67), strict '>' so the first writer of a
 *   maximum wins, table initialised to 0 with back pointers -1 (:50-52; a candidate that does not beat
 *   the initial 0 never writes, so a path may be truncated at the front);
 *   one traceback per r from (n-1, r) counting distinct colours (:70-102).
 * The best_r angle rule (:116-136) is floating point and stays with the caller.
 * Pinned against the reference itself: tests/golden tiny_*_p1_*.dgd (hap_in.* -> hap_out.*) and
 * mhc4_chm13_hapin.npz + expected.json (colors_by_r, path digests) — tests/test_dp_haploid_cpu.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* paths: for every r the traceback path in source->sink order (the reference reverses it at :153),
 * concatenated; path_off[R+2]; path_val has room for path_cap entries. Returns 0, -1 bad args,
 * -2 allocation, -3 path_cap too small. */
int dgo_dp_haploid(int32_t n, const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                   const int64_t* col_off, const int32_t* col_val, int32_t n_colours, int32_t R,
                   int32_t* colours_by_r, int64_t* path_off, int32_t* path_val, int64_t path_cap) {
    if (n <= 0 || R < 0) return -1;
    const size_t S = (size_t)R + 1, N = (size_t)n * S;
    int32_t* dp = (int32_t*)calloc(N, sizeof(int32_t));
    int32_t* bv = (int32_t*)malloc(N * sizeof(int32_t));
    int32_t* br = (int32_t*)malloc(N * sizeof(int32_t));
    uint8_t* seen = (uint8_t*)malloc((size_t)(n_colours > 0 ? n_colours : 1));
    if (!dp || !bv || !br || !seen) { free(dp); free(bv); free(br); free(seen); return -2; }
    memset(bv, 0xFF, N * sizeof(int32_t));
    memset(br, 0xFF, N * sizeof(int32_t));
    for (int32_t u = 0; u < n; ++u)                                   /* :55 */
        for (int32_t r = 0; r <= R; ++r)                              /* :56 */
            for (int64_t e = adj_off[u]; e < adj_off[u + 1]; ++e) {   /* :58 */
                const int32_t v = adj_dst[e], w = adj_w[e];
                if (r + w > R) continue;
                const int32_t cand = dp[(size_t)u * S + r] + (int32_t)(col_off[v + 1] - col_off[v]);
                if (cand > dp[(size_t)v * S + r + w]) {               /* :60 (values are >= 0, so the size_t compare is the signed one) */
                    dp[(size_t)v * S + r + w] = cand;
                    bv[(size_t)v * S + r + w] = u;
                    br[(size_t)v * S + r + w] = r;
                }
            }
    int64_t pos = 0;
    int rc = 0;
    for (int32_t r = 0; r <= R && rc == 0; ++r) {                     /* :74-102 */
        memset(seen, 0, (size_t)(n_colours > 0 ? n_colours : 1));
        int32_t distinct = 0, cv = n - 1, cr = r;
        const int64_t start = pos;
        path_off[r] = pos;
        while (cv != -1) {
            for (int64_t c = col_off[cv]; c < col_off[cv + 1]; ++c)
                if (!seen[col_val[c]]) { seen[col_val[c]] = 1; ++distinct; }
            if (pos >= path_cap) { rc = -3; break; }
            path_val[pos++] = cv;
            const int32_t nv = bv[(size_t)cv * S + cr];
            cr = br[(size_t)cv * S + cr];
            cv = nv;
        }
        colours_by_r[r] = distinct;
        for (int64_t a = start, b = pos - 1; a < b; ++a, --b) { int32_t t = path_val[a]; path_val[a] = path_val[b]; path_val[b] = t; }
    }
    path_off[R + 1] = pos;
    free(dp); free(bv); free(br); free(seen);
    return rc;
}
