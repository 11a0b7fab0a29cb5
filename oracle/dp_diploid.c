/* oracle/dp_diploid.c — TEST INFRASTRUCTURE ONLY (CPU restatement; never shipped, never
 * called by the product path).
 *
 * Plain-C restatement of the reference's diploid recombination-constrained DP:
 *   Approximator::diploid_dp_approximation_solver, /root/reference/src/approximator.cpp:362-785
 *     colour split into hom/het sorted lists ........ :431-453
 *     per-level pair-score table (delta, s_het) ..... :579-624  (set kernels :269-311)
 *     relaxation with (value, smaller i, smaller j) tie-break, payload copy, recombination-edge
 *     lists, both lists extended on the transition into the sink level ... :627-701
 *     answer = cell (r=R,0,0) of the last level; edge lists materialised oldest-first ... :757-785
 * The loop nest (r, i, j, e1, e2) is kept in the reference's order and the comparison is the
 * reference's, so the winner of every cell is the same cell the serial (-t1) reference picks.
 * Only the bookkeeping differs: recombination-edge chains live in one index-linked pool with a
 * mark/compact pass every 1000 levels (reference: per-thread deques + compact_p{1,2}_pool :476-530).
 *
 * Input = the levelized ExpandedGraph (src/ExpandedGraph.hpp:16-26 after
 * strict_bfs_levelize_and_reorder :269-409) in flat form: vertices are numbered in (level,id)
 * order so level l owns ids [level_off[l], level_off[l+1]) and a vertex's position in its level
 * is id - level_off[l]; every edge goes from level l to level l+1.
 *
 * Parity pin: tests/test_dp_diploid_cpu.py checks this file against the reference binary's own
 * per-level DP checksums, sink value, s_het and edge lists (oracle/ref_driver dumps).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define DGO_NEG_INF (INT32_MIN / 4) /* approximator.cpp:413 */

typedef struct { int32_t from, to, prev; } edge_node_t;   /* EdgeNode :381-388, prev = pool index or -1 */

typedef struct {
    int32_t pred_i, pred_j, value, s_het;                 /* dp_entry :390-410 */
    int32_t p1_tail, p2_tail;
} cell_t;

typedef struct { edge_node_t* a; int64_t n, m; } pool_t;

static int32_t pool_push(pool_t* p, int32_t from, int32_t to, int32_t prev) {
    if (p->n == p->m) {
        p->m = p->m ? p->m * 2 : 1024;
        p->a = (edge_node_t*)realloc(p->a, (size_t)p->m * sizeof(edge_node_t));
    }
    p->a[p->n].from = from; p->a[p->n].to = to; p->a[p->n].prev = prev;
    return (int32_t)(p->n++);
}

/* keep only nodes reachable from the live tails (reference: compact_p1_pool/compact_p2_pool) */
static int32_t pool_clone(pool_t* old, pool_t* neu, int32_t* remap, int32_t node) {
    /* iterative: walk down until a remapped node or the root, then rebuild upwards */
    int32_t stack[4096]; int sp = 0; int32_t cur = node;
    while (cur >= 0 && remap[cur] < 0) { stack[sp++] = cur; cur = old->a[cur].prev; if (sp >= 4096) break; }
    int32_t base = cur >= 0 ? remap[cur] : -1;
    while (sp > 0) {
        int32_t x = stack[--sp];
        base = pool_push(neu, old->a[x].from, old->a[x].to, base);
        remap[x] = base;
    }
    return base;
}

static void pool_compact(pool_t* p, cell_t* buf, size_t ncell, int which) {
    pool_t neu = {0, 0, 0};
    int32_t* remap = (int32_t*)malloc((size_t)(p->n ? p->n : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < p->n; ++i) remap[i] = -1;
    for (size_t t = 0; t < ncell; ++t) {
        int32_t* tail = which == 1 ? &buf[t].p1_tail : &buf[t].p2_tail;
        if (*tail >= 0) *tail = pool_clone(p, &neu, remap, *tail);
    }
    free(remap); free(p->a);
    *p = neu;
}

/* |(A u B) n (C u D)| on sorted lists, approximator.cpp:269-288 */
static int inter_union2x2(const int32_t* A, int na, const int32_t* B, int nb,
                          const int32_t* C, int nc, const int32_t* D, int nd) {
    int i = 0, j = 0, k = 0, m = 0, cnt = 0;
    while (i < na || j < nb || k < nc || m < nd) {
        int32_t x = INT32_MAX;
        if (i < na && A[i] < x) x = A[i];
        if (j < nb && B[j] < x) x = B[j];
        if (k < nc && C[k] < x) x = C[k];
        if (m < nd && D[m] < x) x = D[m];
        int l = 0, r = 0;
        while (i < na && A[i] == x) { l = 1; ++i; }
        while (j < nb && B[j] == x) { l = 1; ++j; }
        while (k < nc && C[k] == x) { r = 1; ++k; }
        while (m < nd && D[m] == x) { r = 1; ++m; }
        if (l && r) ++cnt;
    }
    return cnt;
}

/* |(E u F) symdiff (G u H)| on sorted lists, approximator.cpp:292-311 */
static int symd_union2x2(const int32_t* A, int na, const int32_t* B, int nb,
                         const int32_t* C, int nc, const int32_t* D, int nd) {
    int i = 0, j = 0, k = 0, m = 0, cnt = 0;
    while (i < na || j < nb || k < nc || m < nd) {
        int32_t x = INT32_MAX;
        if (i < na && A[i] < x) x = A[i];
        if (j < nb && B[j] < x) x = B[j];
        if (k < nc && C[k] < x) x = C[k];
        if (m < nd && D[m] < x) x = D[m];
        int l = 0, r = 0;
        while (i < na && A[i] == x) { l = 1; ++i; }
        while (j < nb && B[j] == x) { l = 1; ++j; }
        while (k < nc && C[k] == x) { r = 1; ++k; }
        while (m < nd && D[m] == x) { r = 1; ++m; }
        if (l ^ r) ++cnt;
    }
    return cnt;
}

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return x < y ? -1 : x > y;
}

/* same fold as dg_ref_level_done in oracle/ref_hook.h */
static void level_fold(const cell_t* buf, size_t n, uint64_t* sum, uint64_t* live_out) {
    uint64_t h = 1469598103934665603ull, live = 0;
    for (size_t t = 0; t < n; ++t) {
        if (buf[t].value == DGO_NEG_INF) continue;
        ++live;
        uint64_t x = (uint64_t)t * 0x9E3779B97F4A7C15ull;
        x ^= (uint64_t)(uint32_t)buf[t].value * 0xC2B2AE3D27D4EB4Full;
        x ^= ((uint64_t)(uint32_t)buf[t].pred_i << 32 | (uint32_t)buf[t].pred_j) * 0x165667B19E3779F9ull;
        x ^= x >> 29;
        h += x * 0xBF58476D1CE4E5B9ull;
    }
    *sum = h; *live_out = live;
}

/* Returns 0 on success, <0 on malformed input.  p1_edges/p2_edges: capacity 2*(R+2) int32 each,
 * filled with (from,to) pairs oldest-first; n_p1 and n_p2 receive the number of pairs.
 * level_checksum/level_live (nullable): [n_levels], entry l = fold of the DP layer of level l (l>=1). */
int dgo_dp_diploid(int32_t n_levels, const int32_t* level_off,
                   const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                   const int64_t* col_off, const int32_t* col_val,
                   const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                   int32_t* sink_value, int32_t* sink_s_het,
                   int32_t* p1_edges, int32_t* n_p1, int32_t* p2_edges, int32_t* n_p2,
                   uint64_t* level_checksum, uint64_t* level_live) {
    if (n_levels < 1 || R < 0) return -1;
    if (level_off[1] - level_off[0] != 1) return -2;       /* single source on level 0 (:375-377) */
    const int32_t V = level_off[n_levels];
    const int L = n_levels;

    /* hom / het sorted, de-duplicated colour lists per vertex (:431-453) */
    int64_t ncol_tot = col_off[V];
    int32_t* hom = (int32_t*)malloc((size_t)(ncol_tot + 1) * sizeof(int32_t));
    int32_t* het = (int32_t*)malloc((size_t)(ncol_tot + 1) * sizeof(int32_t));
    int64_t* hom_off = (int64_t*)malloc((size_t)(V + 1) * sizeof(int64_t));
    int64_t* het_off = (int64_t*)malloc((size_t)(V + 1) * sizeof(int64_t));
    int64_t nh = 0, nt = 0;
    for (int32_t v = 0; v < V; ++v) {
        hom_off[v] = nh; het_off[v] = nt;
        for (int64_t c = col_off[v]; c < col_off[v + 1]; ++c) {
            int32_t col = col_val[c];
            if (col < 0 || col >= n_colours) { free(hom); free(het); free(hom_off); free(het_off); return -3; }
            if (colour_is_hom[col] == 1) hom[nh++] = col; else het[nt++] = col;
        }
        int64_t a = hom_off[v], b = het_off[v];
        qsort(hom + a, (size_t)(nh - a), sizeof(int32_t), cmp_i32);
        qsort(het + b, (size_t)(nt - b), sizeof(int32_t), cmp_i32);
        int64_t o = a;
        for (int64_t x = a; x < nh; ++x) if (x == a || hom[x] != hom[x - 1]) hom[o++] = hom[x];
        nh = o;
        o = b;
        for (int64_t x = b; x < nt; ++x) if (x == b || het[x] != het[x - 1]) het[o++] = het[x];
        nt = o;
    }
    hom_off[V] = nh; het_off[V] = nt;

    pool_t P1 = {0, 0, 0}, P2 = {0, 0, 0};
    size_t cap_cur = (size_t)(R + 1), cap_next = 0;
    cell_t* cur = (cell_t*)malloc(cap_cur * sizeof(cell_t));
    cell_t* next = NULL;
    for (int r = 0; r <= R; ++r) {                         /* dp_cur.assign(R+1, dp_entry(0,0)) :535 */
        cur[r].pred_i = INT_MAX; cur[r].pred_j = INT_MAX; cur[r].value = 0; cur[r].s_het = 0;
        cur[r].p1_tail = -1; cur[r].p2_tail = -1;
    }
    int32_t* deltas = NULL; int32_t* shets = NULL; size_t cap_d = 0;
    int64_t* base = NULL; size_t cap_b = 0;

    for (int l = 0; l + 1 < L; ++l) {
        const int32_t v0 = level_off[l], k = level_off[l + 1] - v0;
        const int32_t w0 = level_off[l + 1], k2 = level_off[l + 2] - w0;
        const size_t szN = (size_t)(R + 1) * k2 * k2;
        if (szN > cap_next) { cap_next = szN; next = (cell_t*)realloc(next, cap_next * sizeof(cell_t)); }
        for (size_t t = 0; t < szN; ++t) {                 /* reset :565-576 */
            next[t].value = DGO_NEG_INF; next[t].s_het = 0;
            next[t].pred_i = INT_MAX; next[t].pred_j = INT_MAX;
            next[t].p1_tail = -1; next[t].p2_tail = -1;
        }
        /* counts + prefix (:579-601) */
        size_t kk = (size_t)k * k;
        if (kk + 1 > cap_b) { cap_b = kk + 1; base = (int64_t*)realloc(base, cap_b * sizeof(int64_t)); }
        int64_t total = 0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                base[(size_t)i * k + j] = total;
                total += (adj_off[v0 + i + 1] - adj_off[v0 + i]) * (adj_off[v0 + j + 1] - adj_off[v0 + j]);
            }
        base[kk] = total;
        if ((size_t)total > cap_d) {
            cap_d = (size_t)total;
            deltas = (int32_t*)realloc(deltas, cap_d * sizeof(int32_t));
            shets = (int32_t*)realloc(shets, cap_d * sizeof(int32_t));
        }
        /* fill (:604-624) */
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                int32_t u1 = v0 + i, v1 = v0 + j;
                int64_t out = base[(size_t)i * k + j];
                for (int64_t e1 = adj_off[u1]; e1 < adj_off[u1 + 1]; ++e1)
                    for (int64_t e2 = adj_off[v1]; e2 < adj_off[v1 + 1]; ++e2) {
                        int32_t u2 = adj_dst[e1], v2 = adj_dst[e2];
                        int inter = inter_union2x2(hom + hom_off[u1], (int)(hom_off[u1 + 1] - hom_off[u1]),
                                                   hom + hom_off[v1], (int)(hom_off[v1 + 1] - hom_off[v1]),
                                                   hom + hom_off[u2], (int)(hom_off[u2 + 1] - hom_off[u2]),
                                                   hom + hom_off[v2], (int)(hom_off[v2 + 1] - hom_off[v2]));
                        int symd = symd_union2x2(het + het_off[u1], (int)(het_off[u1 + 1] - het_off[u1]),
                                                 het + het_off[v1], (int)(het_off[v1 + 1] - het_off[v1]),
                                                 het + het_off[u2], (int)(het_off[u2 + 1] - het_off[u2]),
                                                 het + het_off[v2], (int)(het_off[v2 + 1] - het_off[v2]));
                        shets[out] = symd;
                        deltas[out] = inter + symd;
                        ++out;
                    }
            }
        /* relaxation (:627-701), serial order r, i, j, e1, e2 */
        const int into_sink = (l + 1 == L - 1);
        for (int r = 0; r <= R; ++r)
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    const cell_t src = cur[((size_t)r * k + i) * k + j];
                    if (src.value == DGO_NEG_INF) continue;
                    int32_t u1 = v0 + i, v1 = v0 + j;
                    int64_t idx = base[(size_t)i * k + j];
                    for (int64_t e1 = adj_off[u1]; e1 < adj_off[u1 + 1]; ++e1) {
                        int32_t u2 = adj_dst[e1]; int wu = adj_w[e1];
                        int iu2 = u2 - w0;
                        if (iu2 < 0 || iu2 >= k2) return -4;
                        for (int64_t e2 = adj_off[v1]; e2 < adj_off[v1 + 1]; ++e2, ++idx) {
                            int32_t v2 = adj_dst[e2]; int wv = adj_w[e2];
                            int jv2 = v2 - w0;
                            if (jv2 < 0 || jv2 >= k2) return -4;
                            int r2 = r + wu + wv;
                            if (r2 > R) continue;
                            cell_t* dst = &next[((size_t)r2 * k2 + iu2) * k2 + jv2];
                            int cand = src.value + deltas[idx];
                            if (cand > dst->value ||
                                (cand == dst->value && i < dst->pred_i) ||
                                (cand == dst->value && i == dst->pred_i && j < dst->pred_j)) {
                                dst->value = cand;
                                dst->s_het = src.s_het + shets[idx];
                                dst->pred_i = i; dst->pred_j = j;
                                dst->p1_tail = src.p1_tail; dst->p2_tail = src.p2_tail;
                                if (wu > 0) dst->p1_tail = pool_push(&P1, u1, u2, dst->p1_tail);
                                if (wv > 0) dst->p2_tail = pool_push(&P2, v1, v2, dst->p2_tail);
                                if (into_sink) {
                                    dst->p1_tail = pool_push(&P1, u1, u2, dst->p1_tail);
                                    dst->p2_tail = pool_push(&P2, v1, v2, dst->p2_tail);
                                }
                            }
                        }
                    }
                }
        /* roll (:704-714) */
        { cell_t* t = cur; cur = next; next = t; size_t c = cap_cur; cap_cur = cap_next; cap_next = c; }
        if (level_checksum) level_fold(cur, szN, &level_checksum[l + 1], &level_live[l + 1]);
        if (((l + 1) % 1000) == 0) { pool_compact(&P1, cur, szN, 1); pool_compact(&P2, cur, szN, 2); }
    }

    /* sink cell (r=R,0,0), :774-785 */
    const int32_t ks = level_off[L] - level_off[L - 1];
    const cell_t* sk = &cur[((size_t)R * ks + 0) * ks + 0];
    *sink_value = sk->value; *sink_s_het = sk->s_het;
    int n1 = 0, n2 = 0;
    for (int32_t c = sk->p1_tail; c >= 0; c = P1.a[c].prev) ++n1;
    for (int32_t c = sk->p2_tail; c >= 0; c = P2.a[c].prev) ++n2;
    int rc = 0;
    if (n1 > R + 2 || n2 > R + 2) rc = -5;
    else {
        int t = n1;
        for (int32_t c = sk->p1_tail; c >= 0; c = P1.a[c].prev) { --t; p1_edges[2 * t] = P1.a[c].from; p1_edges[2 * t + 1] = P1.a[c].to; }
        t = n2;
        for (int32_t c = sk->p2_tail; c >= 0; c = P2.a[c].prev) { --t; p2_edges[2 * t] = P2.a[c].from; p2_edges[2 * t + 1] = P2.a[c].to; }
    }
    *n_p1 = n1; *n_p2 = n2;

    free(P1.a); free(P2.a); free(cur); free(next); free(deltas); free(shets); free(base);
    free(hom); free(het); free(hom_off); free(het_off);
    return rc;
}
