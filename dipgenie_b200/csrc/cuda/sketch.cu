// sketch.cu — B200 minimizer sketch, read spectrum and walk-index join
// (include/dipgenie_cuda.h: dg_sketch_minimizers, dg_sketch_reads, dg_index_walks).
//
// Replaces Solver::compute_hashes (reference src/solver.cpp:366-412), Solver::index_kmers (:277-363), the
// read spectrum Sp_R / kmer_count (:526-555, :711-732) and Solver::compute_anchors (:415-446, :560-576).
//
// How the reference's sequential deque scan becomes a data-parallel kernel (bit-exact):
//   * the k-mer order is lexicographic on the upper-cased canonical ASCII string.  Bytes are replaced by
//     their rank among the byte values that actually occur (order preserving), packed most significant
//     symbol first into a 64*NW-bit integer — for pure ACGT input 2 bits/base, so a 31-mer is one 62-bit
//     word and "string compare" is one integer compare; other alphabets take 4 or 8 bits/symbol (NW=2,4);
//   * the deque pops on '>=' (:316/:388), i.e. the window minimum is the RIGHTMOST smallest k-mer of the
//     w k-mers ending at position i: a plain scan of the window with '<=' replacement;
//   * "emit when the hash differs from the previously emitted hash" is local: the previously emitted hash
//     always equals the hash of the previous window's minimizer, so window i emits iff it is the first
//     window of its sequence (and hash != UINT64_MAX, the reference's initial prev_hash) or
//     hash(min_i) != hash(min_{i-1}).  Hashes are only computed when the minimizer position changes;
//   * MurmurHash3_x64_128(seed 0) h1^h2 (:16-24; algorithm by A. Appleby, public domain) is evaluated on
//     the k ASCII bytes reconstituted from the packed canonical k-mer.
// All sequences of a call (reads or walks) are one concatenated base array cut into tiles of SK_TILE
// window positions; a tile is one CTA: coalesced 16-byte loads into shared memory, k-mer keys rolled per
// thread, window minima from shared memory, block scan for ordered compaction (count pass + emit pass).
// The spectrum is a sort/unique of (hash, read) pairs; the join probes a GPU-resident open-addressing
// hash table hash -> spectrum id.  cub (CUDA toolkit) is used for radix sort and prefix sums only.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "dg_common.cuh"

namespace dg {

constexpr int SK_THREADS = 256;
constexpr int SK_TILE = 2048;                  // window positions per CTA
constexpr uint64_t SK_EMPTY = ~0ull;

struct SketchLut {                             // built on the host from the byte values present
    uint8_t rank[256];                         // raw byte -> rank of its upper-cased value
    uint8_t comp[256];                         // rank -> rank of the complement (misc.cpp:103-115)
    uint8_t inv[256];                          // rank -> upper-cased byte
};

template <int NW>
struct Key {
    uint64_t w[NW];                            // w[0] most significant
};
template <int NW>
__device__ __forceinline__ bool key_less(const Key<NW>& a, const Key<NW>& b) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
    }
    return false;
}
template <int NW>
__device__ __forceinline__ bool key_le(const Key<NW>& a, const Key<NW>& b) { return !key_less<NW>(b, a); }

// key = (key << b | sym) mod 2^K
template <int NW>
__device__ __forceinline__ void key_push_low(Key<NW>& a, int b, uint32_t sym, int K) {
#pragma unroll
    for (int i = 0; i < NW; ++i) a.w[i] = (a.w[i] << b) | (i + 1 < NW ? a.w[i + 1] >> (64 - b) : 0ull);
    a.w[NW - 1] |= sym;
    const int top = 64 * NW - K;               // unused high bits
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int lo = 64 * i;                 // this word covers unused-bit indices [lo, lo+64) counted from the top
        if (top >= lo + 64) a.w[i] = 0;
        else if (top > lo) a.w[i] &= (~0ull) >> (top - lo);
    }
}
// key = key >> b | sym << (K - b)
template <int NW>
__device__ __forceinline__ void key_push_high(Key<NW>& a, int b, uint32_t sym, int K) {
#pragma unroll
    for (int i = NW - 1; i >= 0; --i) a.w[i] = (a.w[i] >> b) | (i > 0 ? a.w[i - 1] << (64 - b) : 0ull);
    const int o = K - b;                       // bit offset from the least significant end; 64 % b == 0, no straddling
    a.w[NW - 1 - o / 64] |= (uint64_t)sym << (o % 64);
}
template <int NW>
__device__ __forceinline__ uint32_t key_sym(const Key<NW>& a, int t, int k, int b) {
    const int o = b * (k - 1 - t);
    return (uint32_t)(a.w[NW - 1 - o / 64] >> (o % 64)) & ((1u << b) - 1u);
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t fmix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
// MurmurHash3_x64_128(bytes, len, seed 0) -> h1 ^ h2 (solver.cpp:16-24)
template <class ByteFn>
__device__ __forceinline__ uint64_t murmur64(ByteFn byte, int len) {
    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = len >> 4;
    for (int i = 0; i < nblocks; ++i) {
        uint64_t k1 = 0, k2 = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) { k1 |= (uint64_t)byte(16 * i + t) << (8 * t); k2 |= (uint64_t)byte(16 * i + 8 + t) << (8 * t); }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int base = nblocks << 4, rem = len & 15;
    uint64_t k1 = 0, k2 = 0;
    for (int t = 8; t < rem; ++t) k2 |= (uint64_t)byte(base + t) << (8 * (t - 8));
    for (int t = 0; t < (rem < 8 ? rem : 8); ++t) k1 |= (uint64_t)byte(base + t) << (8 * t);
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

struct SketchArgs {
    const uint8_t* bases;        // concatenated sequences (16-byte padded allocation)
    uint64_t n_bases;
    const uint64_t* seq_off;     // [n_seq+1]
    uint32_t n_seq;
    int k, w, b;
    const SketchLut* lut;
    uint32_t* tile_count;        // [n_tiles]    (count pass)
    const uint64_t* tile_base;   // [n_tiles]    (emit pass)
    uint64_t* out_hash;          // emissions in sequence order
    uint64_t* out_pos;           // global k-mer start position
    uint32_t* out_seq;           // sequence index
};

__device__ __forceinline__ uint32_t seq_of(const uint64_t* seq_off, uint32_t n_seq, uint64_t p) {
    uint32_t lo = 0, hi = n_seq;                 // last s with seq_off[s] <= p
    while (lo + 1 < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(seq_off + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// Shared memory: sym[NS] | pad | canon[NP] | minpos[NP] | luts | scan
template <int NW, bool EMIT>
__global__ void __launch_bounds__(SK_THREADS) sketch_tile_kernel(const SketchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int k = a.k, w = a.w, b = a.b, K = a.k * a.b;
    const int NP = SK_TILE + w;                  // k-mer positions t0-w .. t0+SK_TILE-1
    const int NS = NP + k - 1;                   // bases
    uint8_t* sym = smem;
    Key<NW>* canon = reinterpret_cast<Key<NW>*>(smem + ((NS + 15) & ~15));
    int32_t* minpos = reinterpret_cast<int32_t*>(canon + NP);
    uint8_t* s_rank = reinterpret_cast<uint8_t*>(minpos + NP);
    uint8_t* s_comp = s_rank + 256;
    uint8_t* s_inv = s_comp + 256;
    uint32_t* s_scan = reinterpret_cast<uint32_t*>(s_inv + 256);   // [SK_THREADS/32 + 1]

    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * SK_TILE;
    const int64_t g0 = t0 - w;                   // global index of sym[0] / canon[0]
    s_rank[tid] = a.lut->rank[tid]; s_comp[tid] = a.lut->comp[tid]; s_inv[tid] = a.lut->inv[tid];
    __syncthreads();

    // ---- bases -> ranks (16-byte coalesced loads) ----
    {
        const int64_t A = (g0 < 0 ? 0 : g0) & ~(int64_t)15;
        const int64_t gend = g0 + NS;
        const int nvec = (int)((gend - A + 15) >> 4);
        for (int j = tid; j < nvec; j += SK_THREADS) {
            const int64_t g = A + 16 * (int64_t)j;
            uint4 v = make_uint4(0, 0, 0, 0);
            if ((uint64_t)g < a.n_bases) v = __ldg(reinterpret_cast<const uint4*>(a.bases + g));
            const uint32_t word[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int64_t x = g + t - g0;
                if (x >= 0 && x < NS) sym[x] = s_rank[(word[t >> 2] >> (8 * (t & 3))) & 0xFFu];
            }
        }
    }
    __syncthreads();

    const int PT = (NP + SK_THREADS - 1) / SK_THREADS;
    const int x0 = tid * PT, x1 = min(NP, x0 + PT);
    // sequence of this thread's first position, advanced as the run crosses boundaries
    uint32_t s = 0;
    uint64_t s_beg = 0, s_end = 0;
    auto locate = [&](int64_t p) {
        if (p < 0 || (uint64_t)p >= a.n_bases) { s_beg = 1; s_end = 0; return; }
        s = seq_of(a.seq_off, a.n_seq, (uint64_t)p);
        s_beg = __ldg(a.seq_off + s); s_end = __ldg(a.seq_off + s + 1);
    };
    auto advance = [&](int64_t p) {              // p increases by one per call
        if (p < 0 || (uint64_t)p >= a.n_bases) { s_beg = 1; s_end = 0; return; }
        if (s_end < s_beg) { locate(p); return; }
        while ((uint64_t)p >= s_end && s + 1 < a.n_seq) { ++s; s_beg = s_end; s_end = __ldg(a.seq_off + s + 1); }
    };

    // ---- canonical k-mer keys ----
    if (x0 < x1) {
        Key<NW> fwd, rev;
        bool have = false;
        locate(g0 + x0);
        for (int x = x0; x < x1; ++x) {
            const int64_t p = g0 + x;
            if (x > x0) advance(p);
            const bool valid = s_end >= s_beg && p >= 0 && (uint64_t)p >= s_beg && (uint64_t)p + k <= s_end;
            Key<NW> c;
            if (valid) {
                if (!have) {
#pragma unroll
                    for (int i = 0; i < NW; ++i) { fwd.w[i] = 0; rev.w[i] = 0; }
                    for (int t = 0; t < k; ++t) {
                        key_push_low<NW>(fwd, b, sym[x + t], K);
                        key_push_low<NW>(rev, b, s_comp[sym[x + k - 1 - t]], K);
                    }
                    have = true;
                } else {
                    key_push_low<NW>(fwd, b, sym[x + k - 1], K);
                    key_push_high<NW>(rev, b, s_comp[sym[x + k - 1]], K);
                }
                c = key_less<NW>(rev, fwd) ? rev : fwd;      // std::min(fwd, rev) (:311/:383)
            } else {
                have = false;
#pragma unroll
                for (int i = 0; i < NW; ++i) c.w[i] = ~0ull;
            }
            canon[x] = c;
        }
    }
    __syncthreads();

    // ---- window minima: rightmost smallest of canon[x-w+1 .. x] ----
    uint32_t flags = 0;                           // bit (x - x0): window x is valid (PT <= 32 is guaranteed by the host)
    uint32_t first = 0;                           // bit: first window of its sequence
    if (x0 < x1) {
        locate(g0 + x0);
        for (int x = x0; x < x1; ++x) {
            const int64_t i = g0 + x;
            if (x > x0) advance(i);
            const bool valid = x >= w - 1 && s_end >= s_beg && i - (w - 1) >= (int64_t)s_beg && (uint64_t)i + k <= s_end;
            int m = -1;
            if (valid) {
                m = x - w + 1;
                Key<NW> best = canon[m];
                for (int y = x - w + 2; y <= x; ++y) {
                    const Key<NW> c = canon[y];
                    if (key_le<NW>(c, best)) { best = c; m = y; }
                }
                flags |= 1u << (x - x0);
                if (i - w < (int64_t)s_beg) first |= 1u << (x - x0);
            }
            minpos[x] = m;
        }
    }
    __syncthreads();

    // ---- emission decisions for windows t0 .. t0+SK_TILE-1 (x >= w) ----
    auto hash_at = [&](int m) -> uint64_t {
        const Key<NW> c = canon[m];
        return murmur64([&](int t) -> uint32_t { return s_inv[key_sym<NW>(c, t, k, b)]; }, k);
    };
    uint32_t emit = 0;
    uint64_t hcache = 0;
    int mcache = -2;
    if (x0 < x1) {
        for (int x = max(x0, w); x < x1; ++x) {
            if (!((flags >> (x - x0)) & 1u)) continue;
            const int m = minpos[x];
            bool e;
            if ((first >> (x - x0)) & 1u) {
                hcache = hash_at(m); mcache = m;
                e = hcache != ~0ull;                                     // prev_hash starts at UINT64_MAX (:303/:376)
            } else {
                const int mp = minpos[x - 1];
                if (mp == m) e = false;
                else {
                    const uint64_t hp = (mp == mcache) ? hcache : hash_at(mp);
                    hcache = hash_at(m); mcache = m;
                    e = hcache != hp;
                }
            }
            if (e) emit |= 1u << (x - x0);
        }
    }
    // ---- ordered compaction ----
    const uint32_t mine = __popc(emit);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < SK_THREADS / 32; ++i) { const uint32_t v = s_scan[i]; s_scan[i] = run; run += v; }
        s_scan[SK_THREADS / 32] = run;
    }
    __syncthreads();
    if (!EMIT) {
        if (tid == 0) a.tile_count[blockIdx.x] = s_scan[SK_THREADS / 32];
        return;
    }
    uint64_t o = a.tile_base[blockIdx.x] + s_scan[tid >> 5] + (incl - mine);
    if (emit) {
        locate(g0 + x0);
        for (int x = x0; x < x1; ++x) {
            if (x > x0) advance(g0 + x);
            if ((emit >> (x - x0)) & 1u) {
                const int m = minpos[x];
                a.out_hash[o] = (m == mcache) ? hcache : hash_at(m);
                a.out_pos[o] = (uint64_t)(g0 + m);
                a.out_seq[o] = s;
                ++o;
            }
        }
    }
}

// byte values present (upper-cased) -> 256 flags
__global__ void alphabet_kernel(const uint8_t* bases, uint64_t n, unsigned int* present) {
    __shared__ unsigned int f[256];
    f[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = bases[i];
        if (c >= 'a' && c <= 'z') c -= 32;
        f[c] = 1;
    }
    __syncthreads();
    if (f[threadIdx.x]) present[threadIdx.x] = 1;
}

// ---- spectrum: (hash, read) pairs sorted by hash (stable, reads ascending within a hash) ----
__global__ void spectrum_flag_kernel(const uint64_t* hash, const uint32_t* seq, uint64_t n, uint32_t* head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || hash[i] != hash[i - 1]) ? 1u : 0u;
}
__global__ void spectrum_fill_kernel(const uint64_t* hash, const uint32_t* seq, const uint32_t* head_incl, uint64_t n,
                                     uint64_t* spectrum, uint32_t* count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = head_incl[i] - 1;
    const bool head = i == 0 || hash[i] != hash[i - 1];
    if (head) spectrum[id] = hash[i];
    if (head || seq[i] != seq[i - 1]) atomicAdd(count + id, 1u);      // one per distinct read (std::set, :531; :716-732)
}

// ---- GPU-resident hash table hash -> spectrum id (open addressing, linear probing) ----
__global__ void table_insert_kernel(const uint64_t* spectrum, uint64_t n, uint64_t* keys, uint32_t* vals, uint64_t mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = spectrum[i];
    if (h == SK_EMPTY) return;                    // handled out of band by the probe
    uint64_t slot = h & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), (unsigned long long)SK_EMPTY,
                                                  (unsigned long long)h);
        if (prev == SK_EMPTY || prev == h) { vals[slot] = (uint32_t)i; return; }
        slot = (slot + 1) & mask;
    }
}
__global__ void table_probe_kernel(const uint64_t* hash, const uint32_t* seq, uint64_t n, const uint64_t* keys,
                                   const uint32_t* vals, uint64_t mask, int64_t id_of_empty, uint32_t* sid, uint32_t* flag,
                                   unsigned long long* per_seq_all, unsigned long long* per_seq_hit) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = hash[i];
    int64_t id = -1;
    if (h == SK_EMPTY) id = id_of_empty;
    else {
        uint64_t slot = h & mask;
        for (;;) {
            const uint64_t kx = __ldg(keys + slot);
            if (kx == h) { id = __ldg(vals + slot); break; }
            if (kx == SK_EMPTY) break;
            slot = (slot + 1) & mask;
        }
    }
    sid[i] = (uint32_t)id;
    flag[i] = id >= 0 ? 1u : 0u;
    atomicAdd(per_seq_all + seq[i], 1ull);
    if (id >= 0) atomicAdd(per_seq_hit + seq[i], 1ull);
}

// ---- covered vertices of every hit (solver.cpp:343-357) ----
struct CoverArgs {
    const uint64_t* pos;          // global k-mer start of every emission
    const uint32_t* seq;          // walk of every emission
    const uint32_t* sid;
    const uint32_t* flag;
    const uint32_t* flag_excl;    // exclusive scan of flag = hit index
    uint64_t n;
    const uint64_t* step_start;   // [n_steps+1] global base offset of every walk step
    const int32_t* step_vtx;      // [n_steps]
    const uint64_t* walk_step_off;// [n_walks+1]
    const int32_t* top_order_map;
    int k;
    uint32_t* hit_sid;            // [n_hits]
    uint32_t* hit_nvtx;           // [n_hits]          (count pass)
    const uint64_t* hit_vtx_off;  // [n_hits+1]        (fill pass)
    int32_t* hit_vtx;
};
template <bool FILL>
__global__ void cover_kernel(const CoverArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n || !a.flag[i]) return;
    const uint32_t hit = a.flag_excl[i];
    const uint64_t st = a.pos[i], en = st + (uint64_t)a.k - 1;
    const uint32_t wk = a.seq[i];
    uint64_t lo = a.walk_step_off[wk], hi = a.walk_step_off[wk + 1];
    const uint64_t end_step = hi;
    while (lo + 1 < hi) {                          // last step with step_start <= st
        const uint64_t mid = (lo + hi) >> 1;
        if (a.step_start[mid] <= st) lo = mid; else hi = mid;
    }
    int32_t* out = FILL ? a.hit_vtx + a.hit_vtx_off[hit] : nullptr;
    uint32_t nv = 0;
    const uint64_t first_step = lo;
    for (uint64_t t = lo; t < end_step && a.step_start[t] <= en; ++t) {
        if (a.step_start[t + 1] == a.step_start[t]) continue;       // an empty segment owns no base
        if (a.step_start[t + 1] <= st) continue;
        const int32_t v = a.step_vtx[t];
        bool dup = false;
        if (FILL) { for (uint32_t x = 0; x < nv; ++x) if (out[x] == v) { dup = true; break; } }
        else {
            for (uint64_t u = first_step; u < t && !dup; ++u)
                if (a.step_start[u + 1] != a.step_start[u] && a.step_start[u + 1] > st && a.step_vtx[u] == v) dup = true;
        }
        if (dup) continue;
        if (FILL) out[nv] = v;
        ++nv;
    }
    if (!FILL) { a.hit_nvtx[hit] = nv; a.hit_sid[hit] = a.sid[i]; return; }
    for (uint32_t x = 1; x < nv; ++x) {            // insertion sort by top_order_map (a permutation: no ties)
        const int32_t v = out[x];
        const int32_t key = a.top_order_map[v];
        int y = (int)x - 1;
        while (y >= 0 && a.top_order_map[out[y]] > key) { out[y + 1] = out[y]; --y; }
        out[y + 1] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
struct Emissions {                 // device-resident result of the sketch pass
    DevBuf<uint64_t> hash, pos;
    DevBuf<uint32_t> seq;
    uint64_t n = 0;
};

struct SketchTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~SketchTimer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

static size_t sketch_smem_bytes(int NW, int k, int w) {
    const int NP = SK_TILE + w, NS = NP + k - 1;
    return ((size_t)(NS + 15) & ~(size_t)15) + (size_t)NP * 8 * NW + (size_t)NP * 4 + 768 + (SK_THREADS / 32 + 1) * 4 + 16;
}

template <int NW>
static int run_sketch_pass(dg_ctx* ctx, SketchArgs& a, uint32_t n_tiles, bool emit, int* launches) {
    const size_t smem = sketch_smem_bytes(NW, a.k, a.w);
    if (emit) {
        DG_CUDA(ctx, cudaFuncSetAttribute((const void*)sketch_tile_kernel<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sketch_tile_kernel<NW, true><<<n_tiles, SK_THREADS, smem, ctx->stream>>>(a);
    } else {
        DG_CUDA(ctx, cudaFuncSetAttribute((const void*)sketch_tile_kernel<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sketch_tile_kernel<NW, false><<<n_tiles, SK_THREADS, smem, ctx->stream>>>(a);
    }
    ++*launches;
    DG_CUDA(ctx, cudaGetLastError());
    return DG_OK;
}

// bases/seq_off are HOST buffers; leaves the ordered emissions on the device.
static int sketch_sequences(dg_ctx* ctx, const uint8_t* bases, const uint64_t* seq_off, uint32_t n_seq, int k, int w,
                            Emissions& em, float* kernel_ms, int* launches) {
    cudaStream_t st = ctx->stream;
    em.n = 0;
    if (k < 1 || w < 1 || w > 4096) return fail(ctx, DG_ERR_ARG, "sketch: k=%d w=%d out of range", k, w);
    if ((SK_TILE + w + SK_THREADS - 1) / SK_THREADS > 32) return fail(ctx, DG_ERR_ARG, "sketch: window %d too large", w);
    const uint64_t n_bases = n_seq ? seq_off[n_seq] : 0;
    for (uint32_t s = 0; s < n_seq; ++s)
        if (seq_off[s + 1] < seq_off[s]) return fail(ctx, DG_ERR_ARG, "sketch: sequence offsets are not monotone");
    if (n_bases == 0) return DG_OK;
    DevBuf<uint8_t> d_bases;
    DevBuf<uint64_t> d_off;
    DevBuf<unsigned int> d_present;
    DevBuf<SketchLut> d_lut;
    DG_CUDA(ctx, d_bases.alloc((size_t)n_bases + 32));
    DG_CUDA(ctx, cudaMemsetAsync(d_bases.p + n_bases, 0, 32, st));
    DG_CUDA(ctx, cudaMemcpyAsync(d_bases.p, bases, (size_t)n_bases, cudaMemcpyHostToDevice, st));
    DG_CUDA(ctx, d_off.upload(seq_off, (size_t)n_seq + 1, st));
    DG_CUDA(ctx, d_present.alloc(256));
    DG_CUDA(ctx, cudaMemsetAsync(d_present.p, 0, 256 * sizeof(unsigned int), st));
    SketchTimer tm;
    DG_CUDA(ctx, cudaEventCreate(&tm.e0));
    DG_CUDA(ctx, cudaEventCreate(&tm.e1));
    DG_CUDA(ctx, cudaEventRecord(tm.e0, st));
    alphabet_kernel<<<std::min<uint64_t>(ctx->sm_count * 8, (n_bases + 255) / 256), 256, 0, st>>>(d_bases.p, n_bases, d_present.p);
    ++*launches;
    DG_CUDA(ctx, cudaGetLastError());
    unsigned int present[256];
    DG_CUDA(ctx, cudaMemcpyAsync(present, d_present.p, sizeof present, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    // order-preserving ranks over {A,C,G,T} + whatever else occurs, so that complements always have a rank
    present['A'] = present['C'] = present['G'] = present['T'] = 1;
    SketchLut lut;
    memset(&lut, 0, sizeof lut);
    int rank_of[256], n_sym = 0;
    for (int c = 0; c < 256; ++c) { rank_of[c] = -1; if (present[c]) { rank_of[c] = n_sym; lut.inv[n_sym] = (uint8_t)c; ++n_sym; } }
    for (int c = 0; c < 256; ++c) {
        const int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
        lut.rank[c] = (uint8_t)(rank_of[u] < 0 ? 0 : rank_of[u]);
    }
    for (int r = 0; r < n_sym; ++r) {
        const int c = lut.inv[r];
        const int cc = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
        lut.comp[r] = (uint8_t)rank_of[cc];
    }
    const int b = n_sym <= 4 ? 2 : n_sym <= 16 ? 4 : 8;
    const int K = k * b;
    const int NW = K <= 64 ? 1 : K <= 128 ? 2 : K <= 256 ? 4 : 0;
    if (!NW) return fail(ctx, DG_ERR_ARG, "sketch: k=%d with %d distinct byte values needs %d key bits (max 256)", k, n_sym, K);
    DG_CUDA(ctx, d_lut.upload(&lut, 1, st));

    const uint32_t n_tiles = (uint32_t)((n_bases + SK_TILE - 1) / SK_TILE);
    DevBuf<uint32_t> d_count;
    DevBuf<uint64_t> d_base;
    DG_CUDA(ctx, d_count.alloc(n_tiles));
    DG_CUDA(ctx, d_base.alloc((size_t)n_tiles + 1));
    SketchArgs a;
    memset(&a, 0, sizeof a);
    a.bases = d_bases.p; a.n_bases = n_bases; a.seq_off = d_off.p; a.n_seq = n_seq; a.k = k; a.w = w; a.b = b;
    a.lut = d_lut.p; a.tile_count = d_count.p; a.tile_base = d_base.p;
    int rc = NW == 1 ? run_sketch_pass<1>(ctx, a, n_tiles, false, launches)
           : NW == 2 ? run_sketch_pass<2>(ctx, a, n_tiles, false, launches) : run_sketch_pass<4>(ctx, a, n_tiles, false, launches);
    if (rc) return rc;
    // exclusive scan of the tile counts (as 64-bit)
    std::vector<uint32_t> cnt(n_tiles);
    DG_CUDA(ctx, cudaMemcpyAsync(cnt.data(), d_count.p, (size_t)n_tiles * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<uint64_t> base((size_t)n_tiles + 1);
    base[0] = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) base[t + 1] = base[t] + cnt[t];
    em.n = base[n_tiles];
    DG_CUDA(ctx, cudaMemcpyAsync(d_base.p, base.data(), base.size() * 8, cudaMemcpyHostToDevice, st));
    DG_CUDA(ctx, em.hash.alloc((size_t)em.n));
    DG_CUDA(ctx, em.pos.alloc((size_t)em.n));
    DG_CUDA(ctx, em.seq.alloc((size_t)em.n));
    a.out_hash = em.hash.p; a.out_pos = em.pos.p; a.out_seq = em.seq.p;
    rc = NW == 1 ? run_sketch_pass<1>(ctx, a, n_tiles, true, launches)
       : NW == 2 ? run_sketch_pass<2>(ctx, a, n_tiles, true, launches) : run_sketch_pass<4>(ctx, a, n_tiles, true, launches);
    if (rc) return rc;
    DG_CUDA(ctx, cudaEventRecord(tm.e1, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    DG_CUDA(ctx, cudaEventElapsedTime(&ms, tm.e0, tm.e1));
    if (kernel_ms) *kernel_ms += ms;
    return DG_OK;
}

template <class T>
static T* host_copy(dg_ctx* ctx, const T* dev, size_t n, int* rc) {
    T* h = (T*)malloc(std::max<size_t>(1, n) * sizeof(T));
    if (!h) { *rc = fail(ctx, DG_ERR_NOMEM, "host allocation of %zu bytes failed", n * sizeof(T)); return nullptr; }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(h, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(h); *rc = fail(ctx, DG_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e)); return nullptr; }
    }
    return h;
}

}  // namespace dg

using namespace dg;

#define g_last_stats (ctx->sketch_stats)   // of the last sketch call on this context (diagnostics / bench)

extern "C" {

int dg_sketch_last_stats(dg_ctx* ctx, dg_sketch_stats_t* out) {
    if (!ctx || !out) return DG_ERR_ARG;
    *out = g_last_stats;
    return DG_OK;
}

int dg_sketch_minimizers(dg_ctx* ctx, const uint8_t* bases, const uint64_t* seq_off, uint32_t n_seq, int k, int w,
                         uint64_t* seq_count, uint64_t** hashes, uint64_t** starts) {
    if (!ctx || !seq_off || !hashes || !starts) return DG_ERR_ARG;
    *hashes = nullptr; *starts = nullptr;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    Emissions em;
    memset(&g_last_stats, 0, sizeof g_last_stats);
    float ms = 0.f; int launches = 0;
    int rc = sketch_sequences(ctx, bases, seq_off, n_seq, k, w, em, &ms, &launches);
    if (rc) return rc;
    uint64_t* h = host_copy(ctx, em.hash.p, (size_t)em.n, &rc);
    if (rc) return rc;
    uint64_t* p = host_copy(ctx, em.pos.p, (size_t)em.n, &rc);
    if (rc) { free(h); return rc; }
    uint32_t* s = host_copy(ctx, em.seq.p, (size_t)em.n, &rc);
    if (rc) { free(h); free(p); return rc; }
    if (seq_count) for (uint32_t i = 0; i < n_seq; ++i) seq_count[i] = 0;
    for (uint64_t i = 0; i < em.n; ++i) {
        if (seq_count) ++seq_count[s[i]];
        p[i] -= seq_off[s[i]];                           // start within its own sequence
    }
    free(s);
    *hashes = h; *starts = p;
    g_last_stats.bases = n_seq ? seq_off[n_seq] : 0; g_last_stats.minimizers = em.n; g_last_stats.kernel_ms = ms;
    g_last_stats.launches = launches;
    return DG_OK;
}

int dg_sketch_reads(dg_ctx* ctx, const uint8_t* bases, const uint64_t* read_off, uint32_t n_reads, int k, int w,
                    uint64_t** spectrum, uint32_t** read_count, uint64_t* n_spectrum) {
    if (!ctx || !read_off || !spectrum || !read_count || !n_spectrum) return DG_ERR_ARG;
    *spectrum = nullptr; *read_count = nullptr; *n_spectrum = 0;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Emissions em;
    memset(&g_last_stats, 0, sizeof g_last_stats);
    float ms = 0.f; int launches = 0;
    int rc = sketch_sequences(ctx, bases, read_off, n_reads, k, w, em, &ms, &launches);
    if (rc) return rc;
    g_last_stats.bases = n_reads ? read_off[n_reads] : 0; g_last_stats.minimizers = em.n;
    uint64_t ns = 0;
    if (em.n) {
        const uint64_t n = em.n;
        if (n > 0x7FFFFFFFull) return fail(ctx, DG_ERR_CAPACITY, "dg_sketch_reads: %llu minimizers exceed one sort batch", (unsigned long long)n);
        SketchTimer tm;
        DG_CUDA(ctx, cudaEventCreate(&tm.e0));
        DG_CUDA(ctx, cudaEventCreate(&tm.e1));
        DG_CUDA(ctx, cudaEventRecord(tm.e0, st));
        DevBuf<uint64_t> k2;
        DevBuf<uint32_t> v2, head, head_incl;
        DevBuf<uint8_t> tmp;
        DG_CUDA(ctx, k2.alloc((size_t)n));
        DG_CUDA(ctx, v2.alloc((size_t)n));
        DG_CUDA(ctx, head.alloc((size_t)n));
        DG_CUDA(ctx, head_incl.alloc((size_t)n));
        size_t tb = 0, tb2 = 0;
        DG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, em.hash.p, k2.p, em.seq.p, v2.p, (int)n, 0, 64, st));
        DG_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tb2, head.p, head_incl.p, (int)n, st));
        DG_CUDA(ctx, tmp.alloc(std::max(tb, tb2)));
        DG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, em.hash.p, k2.p, em.seq.p, v2.p, (int)n, 0, 64, st));
        const unsigned grid = (unsigned)((n + 255) / 256);
        spectrum_flag_kernel<<<grid, 256, 0, st>>>(k2.p, v2.p, n, head.p);
        DG_CUDA(ctx, cudaGetLastError());
        DG_CUDA(ctx, cub::DeviceScan::InclusiveSum(tmp.p, tb2, head.p, head_incl.p, (int)n, st));
        uint32_t last = 0;
        DG_CUDA(ctx, cudaMemcpyAsync(&last, head_incl.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(ctx, cudaStreamSynchronize(st));
        ns = last;
        DevBuf<uint64_t> d_sp;
        DevBuf<uint32_t> d_cnt;
        DG_CUDA(ctx, d_sp.alloc((size_t)ns));
        DG_CUDA(ctx, d_cnt.alloc((size_t)ns));
        DG_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, (size_t)ns * 4, st));
        spectrum_fill_kernel<<<grid, 256, 0, st>>>(k2.p, v2.p, head_incl.p, n, d_sp.p, d_cnt.p);
        DG_CUDA(ctx, cudaGetLastError());
        launches += 5;
        DG_CUDA(ctx, cudaEventRecord(tm.e1, st));
        uint64_t* hs = host_copy(ctx, d_sp.p, (size_t)ns, &rc);
        if (rc) return rc;
        uint32_t* hc = host_copy(ctx, d_cnt.p, (size_t)ns, &rc);
        if (rc) { free(hs); return rc; }
        float ms2 = 0.f;
        DG_CUDA(ctx, cudaEventElapsedTime(&ms2, tm.e0, tm.e1));
        ms += ms2;
        *spectrum = hs; *read_count = hc;
    } else {
        *spectrum = (uint64_t*)malloc(8); *read_count = (uint32_t*)malloc(4);
    }
    *n_spectrum = ns;
    g_last_stats.kernel_ms = ms; g_last_stats.launches = launches; g_last_stats.spectrum = ns;
    return DG_OK;
}

int dg_index_walks(dg_ctx* ctx, const uint8_t* seg_bases, const uint64_t* seg_off, uint32_t n_seg, const int32_t* walk_vtx,
                   const uint64_t* walk_off, uint32_t n_walks, const int32_t* top_order_map, int k, int w,
                   const uint64_t* spectrum, uint64_t n_spectrum, uint64_t* n_minimizers, uint64_t** hit_off,
                   uint32_t** hit_sid, uint64_t** hit_vtx_off, int32_t** hit_vtx) {
    if (!ctx || !seg_off || !walk_off || !hit_off || !hit_sid || !hit_vtx_off || !hit_vtx) return DG_ERR_ARG;
    *hit_off = nullptr; *hit_sid = nullptr; *hit_vtx_off = nullptr; *hit_vtx = nullptr;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    memset(&g_last_stats, 0, sizeof g_last_stats);
    // concatenated walk sequences (what index_kmers builds per walk, :283-285) + the base offset of every step
    const uint64_t n_steps = n_walks ? walk_off[n_walks] : 0;
    std::vector<uint64_t> step_start((size_t)n_steps + 1), wseq_off((size_t)n_walks + 1);
    uint64_t tot = 0;
    for (uint32_t h = 0; h < n_walks; ++h) {
        wseq_off[h] = tot;
        for (uint64_t t = walk_off[h]; t < walk_off[h + 1]; ++t) {
            const int32_t v = walk_vtx[t];
            if (v < 0 || (uint32_t)v >= n_seg) return fail(ctx, DG_ERR_ARG, "dg_index_walks: walk %u step %llu names segment %d", h, (unsigned long long)t, v);
            step_start[t] = tot;
            tot += seg_off[v + 1] - seg_off[v];
        }
    }
    wseq_off[n_walks] = tot; step_start[n_steps] = tot;
    std::vector<uint8_t> wb((size_t)tot + 1);
    for (uint64_t t = 0; t < n_steps; ++t) {
        const int32_t v = walk_vtx[t];
        memcpy(wb.data() + step_start[t], seg_bases + seg_off[v], (size_t)(seg_off[v + 1] - seg_off[v]));
    }
    Emissions em;
    float ms = 0.f; int launches = 0;
    int rc = sketch_sequences(ctx, wb.data(), wseq_off.data(), n_walks, k, w, em, &ms, &launches);
    if (rc) return rc;
    g_last_stats.bases = tot; g_last_stats.minimizers = em.n; g_last_stats.spectrum = n_spectrum;
    const uint64_t n = em.n;
    std::vector<unsigned long long> all((size_t)n_walks, 0), hits((size_t)n_walks, 0);
    uint64_t n_hits = 0, n_vtx = 0;
    uint32_t* h_sid = nullptr; uint64_t* h_voff = nullptr; int32_t* h_vtx = nullptr;
    if (n) {
        if (n > 0x7FFFFFFFull) return fail(ctx, DG_ERR_CAPACITY, "dg_index_walks: %llu minimizers exceed one batch", (unsigned long long)n);
        SketchTimer tm;
        DG_CUDA(ctx, cudaEventCreate(&tm.e0));
        DG_CUDA(ctx, cudaEventCreate(&tm.e1));
        // hash table over the spectrum
        uint64_t cap = 64;
        while (cap < 2 * n_spectrum + 2) cap <<= 1;
        int64_t id_of_empty = -1;
        for (uint64_t i = n_spectrum; i-- > 0 && spectrum[i] == SK_EMPTY;) id_of_empty = (int64_t)i;   // ascending: only the last can be ~0
        DevBuf<uint64_t> d_sp, d_keys, d_step_start, d_wso, d_voff;
        DevBuf<uint32_t> d_vals, d_sid, d_flag, d_excl, d_hsid, d_nv;
        DevBuf<unsigned long long> d_all, d_hit;
        DevBuf<int32_t> d_step_vtx, d_tom, d_hvtx;
        DevBuf<uint8_t> tmp;
        DG_CUDA(ctx, d_sp.upload(spectrum, (size_t)n_spectrum, st));
        DG_CUDA(ctx, d_keys.alloc((size_t)cap));
        DG_CUDA(ctx, d_vals.alloc((size_t)cap));
        DG_CUDA(ctx, cudaMemsetAsync(d_keys.p, 0xFF, (size_t)cap * 8, st));
        DG_CUDA(ctx, d_sid.alloc((size_t)n));
        DG_CUDA(ctx, d_flag.alloc((size_t)n));
        DG_CUDA(ctx, d_excl.alloc((size_t)n));
        DG_CUDA(ctx, d_all.alloc((size_t)n_walks));
        DG_CUDA(ctx, d_hit.alloc((size_t)n_walks));
        DG_CUDA(ctx, cudaMemsetAsync(d_all.p, 0, (size_t)n_walks * 8, st));
        DG_CUDA(ctx, cudaMemsetAsync(d_hit.p, 0, (size_t)n_walks * 8, st));
        DG_CUDA(ctx, d_step_start.upload(step_start.data(), step_start.size(), st));
        DG_CUDA(ctx, d_step_vtx.upload(walk_vtx, (size_t)n_steps, st));
        DG_CUDA(ctx, d_wso.upload(walk_off, (size_t)n_walks + 1, st));
        DG_CUDA(ctx, d_tom.upload(top_order_map, (size_t)n_seg, st));
        DG_CUDA(ctx, cudaEventRecord(tm.e0, st));
        if (n_spectrum) {
            table_insert_kernel<<<(unsigned)((n_spectrum + 255) / 256), 256, 0, st>>>(d_sp.p, n_spectrum, d_keys.p, d_vals.p, cap - 1);
            DG_CUDA(ctx, cudaGetLastError());
            ++launches;
        }
        const unsigned grid = (unsigned)((n + 255) / 256);
        table_probe_kernel<<<grid, 256, 0, st>>>(em.hash.p, em.seq.p, n, d_keys.p, d_vals.p, cap - 1, id_of_empty, d_sid.p, d_flag.p,
                                                 d_all.p, d_hit.p);
        DG_CUDA(ctx, cudaGetLastError());
        size_t tb = 0;
        DG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag.p, d_excl.p, (int)n, st));
        DG_CUDA(ctx, tmp.alloc(tb));
        DG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_flag.p, d_excl.p, (int)n, st));
        uint32_t last_excl = 0, last_flag = 0;
        DG_CUDA(ctx, cudaMemcpyAsync(&last_excl, d_excl.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(ctx, cudaMemcpyAsync(&last_flag, d_flag.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all.p, (size_t)n_walks * 8, cudaMemcpyDeviceToHost, st));
        DG_CUDA(ctx, cudaMemcpyAsync(hits.data(), d_hit.p, (size_t)n_walks * 8, cudaMemcpyDeviceToHost, st));
        DG_CUDA(ctx, cudaStreamSynchronize(st));
        n_hits = (uint64_t)last_excl + last_flag;
        launches += 2;
        DG_CUDA(ctx, d_hsid.alloc((size_t)n_hits));
        DG_CUDA(ctx, d_nv.alloc((size_t)n_hits + 1));
        DG_CUDA(ctx, d_voff.alloc((size_t)n_hits + 1));
        CoverArgs ca;
        memset(&ca, 0, sizeof ca);
        ca.pos = em.pos.p; ca.seq = em.seq.p; ca.sid = d_sid.p; ca.flag = d_flag.p; ca.flag_excl = d_excl.p; ca.n = n;
        ca.step_start = d_step_start.p; ca.step_vtx = d_step_vtx.p; ca.walk_step_off = d_wso.p; ca.top_order_map = d_tom.p;
        ca.k = k; ca.hit_sid = d_hsid.p; ca.hit_nvtx = d_nv.p; ca.hit_vtx_off = d_voff.p;
        std::vector<uint64_t> voff((size_t)n_hits + 1, 0);
        if (n_hits) {
            cover_kernel<false><<<grid, 256, 0, st>>>(ca);
            DG_CUDA(ctx, cudaGetLastError());
            std::vector<uint32_t> nv((size_t)n_hits);
            DG_CUDA(ctx, cudaMemcpyAsync(nv.data(), d_nv.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, st));
            DG_CUDA(ctx, cudaStreamSynchronize(st));
            for (uint64_t i = 0; i < n_hits; ++i) voff[i + 1] = voff[i] + nv[i];
            n_vtx = voff[n_hits];
            DG_CUDA(ctx, cudaMemcpyAsync(d_voff.p, voff.data(), voff.size() * 8, cudaMemcpyHostToDevice, st));
            DG_CUDA(ctx, d_hvtx.alloc((size_t)n_vtx));
            ca.hit_vtx = d_hvtx.p;
            cover_kernel<true><<<grid, 256, 0, st>>>(ca);
            DG_CUDA(ctx, cudaGetLastError());
            launches += 2;
        }
        DG_CUDA(ctx, cudaEventRecord(tm.e1, st));
        h_sid = host_copy(ctx, d_hsid.p, (size_t)n_hits, &rc);
        if (rc) return rc;
        h_vtx = host_copy(ctx, d_hvtx.p, (size_t)n_vtx, &rc);
        if (rc) { free(h_sid); return rc; }
        h_voff = (uint64_t*)malloc(((size_t)n_hits + 1) * 8);
        memcpy(h_voff, voff.data(), ((size_t)n_hits + 1) * 8);
        float ms2 = 0.f;
        DG_CUDA(ctx, cudaEventElapsedTime(&ms2, tm.e0, tm.e1));
        ms += ms2;
    } else {
        h_sid = (uint32_t*)malloc(4); h_vtx = (int32_t*)malloc(4); h_voff = (uint64_t*)calloc(1, 8);
    }
    uint64_t* h_off = (uint64_t*)malloc(((size_t)n_walks + 1) * 8);
    h_off[0] = 0;
    for (uint32_t h = 0; h < n_walks; ++h) {
        h_off[h + 1] = h_off[h] + hits[h];
        if (n_minimizers) n_minimizers[h] = all[h];
    }
    *hit_off = h_off; *hit_sid = h_sid; *hit_vtx_off = h_voff; *hit_vtx = h_vtx;
    g_last_stats.kernel_ms = ms; g_last_stats.launches = launches; g_last_stats.hits = n_hits;
    return DG_OK;
}

}  // extern "C"
