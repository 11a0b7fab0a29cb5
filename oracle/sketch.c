/* oracle/sketch.c — TEST INFRASTRUCTURE ONLY (CPU restatement; never shipped, never called by the
 * product path).
 *
 * Plain-C restatement of the reference's minimizer sketch / hit stage:
 *   hash128_to_64_ ............................. /root/reference/src/solver.cpp:16-24
 *     MurmurHash3_x64_128, seed 0 (public-domain algorithm by Austin Appleby; the reference vendors it
 *     as src/MurmurHash3.cpp:255-332) — restated here from the published definition
 *   reverse_strand_ ............................ src/misc.cpp:103-115
 *   Solver::compute_hashes (reads) ............. src/solver.cpp:366-412
 *   Solver::index_kmers (walks) ................ src/solver.cpp:277-363
 *   read spectrum Sp_R + per-hash read count ... src/solver.cpp:526-555, :711-732
 *   Solver::compute_anchors (join) ............. src/solver.cpp:415-446, :560-576
 *
 * Semantics kept exactly:
 *   - the sequence is upper-cased first (toupper); every other byte is kept as is;
 *   - canonical k-mer = min(forward, reverse complement) as ASCII strings; complement maps A<->T, C<->G,
 *     any other byte to itself;
 *   - window minimum over w consecutive k-mers, ties resolved to the RIGHTMOST equal k-mer (the deque pops
 *     on >=, :316/:388);
 *   - from window i = w-1 on, the hash of the window minimum is emitted iff it differs from the previously
 *     emitted hash (prev_hash starts at UINT64_MAX);
 *   - sequences shorter than w+k-1 give nothing.
 *
 * Parity pin: tests/test_sketch_cpu.py compares with the reference's own index_kmers / compute_hashes
 * outputs dumped by oracle/ref_driver (tests/golden/tiny_*.dgd, and the MHC panel when oracle/_ref exists).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

/* MurmurHash3_x64_128(key,len,seed=0) -> h1 ^ h2 (solver.cpp:16-24) */
uint64_t dgo_hash64(const uint8_t* key, int len) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = len / 16;
    for (int i = 0; i < nblocks; ++i) {
        uint64_t k1, k2;
        memcpy(&k1, key + 16 * i, 8);
        memcpy(&k2, key + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* tail = key + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    const int rem = len & 15;
    for (int t = rem; t > 8; --t) k2 ^= (uint64_t)tail[t - 1] << (8 * (t - 9));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int t = rem < 8 ? rem : 8; t > 0; --t) k1 ^= (uint64_t)tail[t - 1] << (8 * (t - 1));
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
static inline uint8_t comp(uint8_t c) {   /* on upper-cased input (misc.cpp:103-115) */
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c; }
}

/* byte t of the canonical k-mer starting at p; rc = 1 when the reverse complement was chosen */
static inline uint8_t canon_byte(const uint8_t* s, int64_t p, int k, int rc, int t) {
    return rc ? comp(s[p + k - 1 - t]) : s[p + t];
}
static int choose_rc(const uint8_t* s, int64_t p, int k) {   /* std::min(fwd, rev): rev only if strictly smaller */
    for (int t = 0; t < k; ++t) {
        uint8_t f = s[p + t], r = comp(s[p + k - 1 - t]);
        if (f != r) return r < f;
    }
    return 0;
}
static int canon_cmp(const uint8_t* s, int k, int64_t p, int rp, int64_t q, int rq) {
    for (int t = 0; t < k; ++t) {
        uint8_t a = canon_byte(s, p, k, rp, t), b = canon_byte(s, q, k, rq, t);
        if (a != b) return a < b ? -1 : 1;
    }
    return 0;
}

/* Minimizer scan over an upper-cased sequence s[0..n).  For every emitted minimizer calls
 * cb(ctx, hash, start_position).  Mirrors the deque loop of solver.cpp:307-361 / :379-409. */
typedef void (*emit_fn)(void* ctx, uint64_t hash, int64_t start);
static void scan_minimizers(const uint8_t* s, int64_t n, int k, int w, emit_fn cb, void* ctx) {
    if (n < (int64_t)w + k - 1) return;
    const int64_t nk = n - k + 1;
    int64_t* dq = (int64_t*)malloc((size_t)(w + 2) * sizeof(int64_t));   /* positions, ring of capacity w+1 */
    uint8_t* rcf = (uint8_t*)malloc((size_t)nk);
    int head = 0, cnt = 0;
    const int cap = w + 2;
    uint64_t prev = UINT64_MAX;
    uint8_t buf[256];
    int64_t last_pos = -1; uint64_t last_hash = 0;
    for (int64_t i = 0; i < nk; ++i) {
        rcf[i] = (uint8_t)choose_rc(s, i, k);
        while (cnt > 0) {
            int64_t b = dq[(head + cnt - 1) % cap];
            if (canon_cmp(s, k, b, rcf[b], i, rcf[i]) >= 0) --cnt; else break;   /* back >= current: pop */
        }
        dq[(head + cnt) % cap] = i; ++cnt;
        if (cnt > 0 && dq[head] <= i - w) { head = (head + 1) % cap; --cnt; }
        if (i >= w - 1) {
            int64_t p = dq[head];
            uint64_t h;
            if (p == last_pos) h = last_hash;
            else {
                for (int t = 0; t < k; ++t) buf[t] = canon_byte(s, p, k, rcf[p], t);
                h = dgo_hash64(buf, k);
                last_pos = p; last_hash = h;
            }
            if (h != prev) { prev = h; cb(ctx, h, p); }
        }
    }
    free(dq); free(rcf);
}

typedef struct { uint64_t* a; size_t n, m; } u64vec;
static void u64_push(u64vec* v, uint64_t x) {
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (uint64_t*)realloc(v->a, v->m * 8); }
    v->a[v->n++] = x;
}
static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}
static void emit_hash_only(void* ctx, uint64_t h, int64_t start) { (void)start; u64_push((u64vec*)ctx, h); }

/* Read sketch + spectrum (solver.cpp:526-555) + per-hash read multiplicity (:711-732).
 * bases: concatenated reads (any case), read i = bases[read_off[i] .. read_off[i+1]).
 * Outputs (malloc'ed, caller frees with free()): spectrum = distinct hashes ascending (id = index),
 * read_count[id] = number of reads whose hash set contains it.  Optionally per-read hash sets
 * (ascending within a read, like std::set iteration) via per_read_off/per_read_val (nullable). */
int dgo_sketch_reads(const uint8_t* bases, const uint64_t* read_off, uint32_t n_reads, int k, int w,
                     uint64_t** spectrum, uint32_t** read_count, uint64_t* n_spectrum,
                     uint64_t** per_read_off, uint64_t** per_read_val) {
    if (k < 1 || k > 255 || w < 1) return -1;
    u64vec all = {0, 0, 0};
    uint64_t* pro = per_read_off ? (uint64_t*)malloc(((size_t)n_reads + 1) * 8) : NULL;
    if (pro) pro[0] = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        const int64_t n = (int64_t)(read_off[r + 1] - read_off[r]);
        uint8_t* s = (uint8_t*)malloc((size_t)n + 1);
        for (int64_t t = 0; t < n; ++t) s[t] = up(bases[read_off[r] + t]);
        size_t before = all.n;
        scan_minimizers(s, n, k, w, emit_hash_only, &all);
        free(s);
        /* per-read set: sort + unique the tail */
        qsort(all.a + before, all.n - before, 8, cmp_u64);
        size_t o = before;
        for (size_t x = before; x < all.n; ++x) if (x == before || all.a[x] != all.a[x - 1]) all.a[o++] = all.a[x];
        all.n = o;
        if (pro) pro[r + 1] = all.n;
    }
    if (per_read_val) {
        *per_read_val = (uint64_t*)malloc((all.n ? all.n : 1) * 8);
        memcpy(*per_read_val, all.a, all.n * 8);
        *per_read_off = pro;
    }
    qsort(all.a, all.n, 8, cmp_u64);
    uint64_t* sp = (uint64_t*)malloc((all.n ? all.n : 1) * 8);
    uint32_t* rc = (uint32_t*)malloc((all.n ? all.n : 1) * 4);
    size_t ns = 0;
    for (size_t x = 0; x < all.n; ++x) {
        if (x == 0 || all.a[x] != all.a[x - 1]) { sp[ns] = all.a[x]; rc[ns] = 1; ++ns; }
        else ++rc[ns - 1];
    }
    free(all.a);
    *spectrum = sp; *read_count = rc; *n_spectrum = ns;
    return 0;
}

typedef struct {
    u64vec hash, start;
} walk_emit_t;
static void emit_walk(void* ctx, uint64_t h, int64_t start) {
    walk_emit_t* e = (walk_emit_t*)ctx;
    u64_push(&e->hash, h); u64_push(&e->start, (uint64_t)start);
}

static int64_t find_spectrum(const uint64_t* sp, uint64_t n, uint64_t h) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (sp[mid] < h) lo = mid + 1; else hi = mid; }
    return (lo < n && sp[lo] == h) ? (int64_t)lo : -1;
}

/* Walk index (index_kmers) joined with the read spectrum (compute_anchors):
 * for every walk, in walk order, the minimizers whose hash is in `spectrum`, each with the unique vertices
 * covered by bases [start, start+k) ordered by top_order_map (:343-357).
 * Outputs (malloc'ed): n_minimizers[n_walks] (all minimizers of the walk, before the join: the
 * "Number of Minimizers" log line, :474), hit_off[n_walks+1], hit_sid[n_hits] (spectrum id),
 * hit_vtx_off[n_hits+1], hit_vtx[].  all_hash/all_off (nullable): every minimizer hash of every walk. */
int dgo_index_walks(const uint8_t* seg_bases, const uint64_t* seg_off, uint32_t n_seg,
                    const int32_t* walk_vtx, const uint64_t* walk_off, uint32_t n_walks,
                    const int32_t* top_order_map, int k, int w,
                    const uint64_t* spectrum, uint64_t n_spectrum,
                    uint64_t* n_minimizers, uint64_t** hit_off, uint32_t** hit_sid,
                    uint64_t** hit_vtx_off, int32_t** hit_vtx,
                    uint64_t** all_off, uint64_t** all_hash) {
    if (k < 1 || k > 255 || w < 1) return -1;
    (void)n_seg;
    uint64_t* hoff = (uint64_t*)malloc(((size_t)n_walks + 1) * 8);
    u64vec sid = {0, 0, 0}, voff = {0, 0, 0}, vtx = {0, 0, 0}, allh = {0, 0, 0};
    uint64_t* aoff = all_off ? (uint64_t*)malloc(((size_t)n_walks + 1) * 8) : NULL;
    hoff[0] = 0; u64_push(&voff, 0);
    if (aoff) aoff[0] = 0;
    for (uint32_t h = 0; h < n_walks; ++h) {
        const uint64_t p0 = walk_off[h], p1 = walk_off[h + 1];
        int64_t n = 0;
        for (uint64_t t = p0; t < p1; ++t) n += (int64_t)(seg_off[walk_vtx[t] + 1] - seg_off[walk_vtx[t]]);
        uint8_t* s = (uint8_t*)malloc((size_t)n + 1);
        int64_t* vstart = (int64_t*)malloc(((size_t)(p1 - p0) + 1) * sizeof(int64_t));   /* base offset of each walk step */
        int64_t pos = 0;
        for (uint64_t t = p0; t < p1; ++t) {
            const int32_t v = walk_vtx[t];
            vstart[t - p0] = pos;
            for (uint64_t b = seg_off[v]; b < seg_off[v + 1]; ++b) s[pos++] = up(seg_bases[b]);
        }
        vstart[p1 - p0] = pos;
        walk_emit_t e; memset(&e, 0, sizeof e);
        scan_minimizers(s, n, k, w, emit_walk, &e);
        n_minimizers[h] = e.hash.n;
        for (size_t m = 0; m < e.hash.n; ++m) {
            if (all_hash) u64_push(&allh, e.hash.a[m]);
            const int64_t id = find_spectrum(spectrum, n_spectrum, e.hash.a[m]);
            if (id < 0) continue;
            /* unique vertices under bases [start, start+k), first-appearance order, then sorted by top_order_map */
            const int64_t st = (int64_t)e.start.a[m], en = st + k - 1;
            int64_t lo = 0, hi = (int64_t)(p1 - p0);
            while (lo + 1 < hi) { int64_t mid = (lo + hi) / 2; if (vstart[mid] <= st) lo = mid; else hi = mid; }
            /* steps with empty sequence never own a base: skip them like idx_vtx_map does (:294-300) */
            int32_t tmp[256]; int nt = 0;
            for (int64_t t = lo; t < (int64_t)(p1 - p0) && vstart[t] <= en; ++t) {
                if (vstart[t + 1] == vstart[t]) continue;
                if (vstart[t + 1] <= st) continue;
                const int32_t v = walk_vtx[p0 + t];
                int dup = 0;
                for (int x = 0; x < nt; ++x) if (tmp[x] == v) { dup = 1; break; }
                if (!dup && nt < 256) tmp[nt++] = v;
            }
            for (int a = 1; a < nt; ++a) {           /* insertion sort by top_order_map (a permutation: no ties) */
                int32_t v = tmp[a]; int b = a - 1;
                while (b >= 0 && top_order_map[tmp[b]] > top_order_map[v]) { tmp[b + 1] = tmp[b]; --b; }
                tmp[b + 1] = v;
            }
            u64_push(&sid, (uint64_t)id);
            for (int x = 0; x < nt; ++x) u64_push(&vtx, (uint64_t)(uint32_t)tmp[x]);
            u64_push(&voff, vtx.n);
        }
        hoff[h + 1] = sid.n;
        if (aoff) aoff[h + 1] = allh.n;
        free(e.hash.a); free(e.start.a); free(s); free(vstart);
    }
    uint32_t* sid32 = (uint32_t*)malloc((sid.n ? sid.n : 1) * 4);
    for (size_t x = 0; x < sid.n; ++x) sid32[x] = (uint32_t)sid.a[x];
    int32_t* v32 = (int32_t*)malloc((vtx.n ? vtx.n : 1) * 4);
    for (size_t x = 0; x < vtx.n; ++x) v32[x] = (int32_t)vtx.a[x];
    *hit_off = hoff; *hit_sid = sid32; *hit_vtx_off = voff.a; *hit_vtx = v32;
    free(sid.a); free(vtx.a);
    if (all_off) { *all_off = aoff; *all_hash = allh.a; } else free(allh.a);
    return 0;
}

void dgo_free(void* p) { free(p); }
