// anchors.cpp — host glue between the device sketch/join stage and the graph expansion: the second half
// of Solver::compute_and_classify_anchors (reference src/solver.cpp:560-887).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "dgh.h"

namespace dgh {

void build_anchors(const Panel& p, const SketchResult& s, float threshold, int threads, Anchors& a) {
    a = Anchors();
    const int H = (int)p.paths.size();
    const int64_t S = (int64_t)s.spectrum.size();
    a.n_ids = (int32_t)S; a.n_walks = H;

    // Anchor_hits[id][h] in walk order (:560-576): bucket the hits by (id, walk)
    const int64_t n_hits = (int64_t)s.hit_sid.size();
    std::vector<int64_t> bucket_off((size_t)S * H + 1, 0);
    std::vector<int32_t> hit_walk((size_t)n_hits);
    for (int h = 0; h < H; ++h)
        for (uint64_t x = s.hit_off[h]; x < s.hit_off[h + 1]; ++x) {
            hit_walk[x] = h;
            ++bucket_off[(size_t)s.hit_sid[x] * H + h + 1];
        }
    for (size_t i = 0; i < (size_t)S * H; ++i) bucket_off[i + 1] += bucket_off[i];
    std::vector<int64_t> bucket((size_t)n_hits);
    {
        std::vector<int64_t> fill(bucket_off.begin(), bucket_off.end() - 1);
        for (int64_t x = 0; x < n_hits; ++x) bucket[(size_t)fill[(size_t)s.hit_sid[x] * H + hit_walk[x]]++] = x;
    }

    // panel-sharing filter (:590-638) + occurrence order (:641-663)
    a.occ_off.assign((size_t)S * H + 1, 0);
    a.vtx_off.assign(1, 0);
    a.anchors_per_walk.assign(H, 0);
    std::vector<std::vector<int64_t>> kept(H);     // per walk: hit indices of the current id, in the reference's order
    // The reference groups the occurrences of an id in a std::map keyed by the TEXT "v1_v2_..._" of their vertex lists (:596-611):
    // groups come out in the lexicographic order of those strings, occurrences inside a group in insertion order (walk, then
    // position).  The same order without building a string or a map node per occurrence: a stable sort of the occurrences by
    // a comparator that is the string comparison — the first differing vertex decides, by its decimal digits followed by
    // '_' (which sorts above every digit), and a list that is a prefix of the other comes first.
    auto token_less = [](int32_t x, int32_t y) {          // "x_" < "y_" as strings, x != y
        char a[16], b[16];
        const int la = snprintf(a, sizeof a, "%d_", x), lb = snprintf(b, sizeof b, "%d_", y);
        const int c = memcmp(a, b, (size_t)std::min(la, lb));
        return c != 0 ? c < 0 : la < lb;
    };
    auto key_cmp = [&](int64_t A, int64_t B) {            // <0, 0, >0
        const uint64_t a0 = s.hit_vtx_off[A], a1 = s.hit_vtx_off[A + 1], b0 = s.hit_vtx_off[B], b1 = s.hit_vtx_off[B + 1];
        const uint64_t n = std::min(a1 - a0, b1 - b0);
        for (uint64_t t = 0; t < n; ++t) {
            const int32_t x = s.hit_vtx[a0 + t], y = s.hit_vtx[b0 + t];
            if (x != y) return token_less(x, y) ? -1 : 1;
        }
        return (a1 - a0) < (b1 - b0) ? -1 : ((a1 - a0) > (b1 - b0) ? 1 : 0);
    };
    std::vector<int64_t> occ;                              // occurrences of the current id, grouped
    for (int64_t id = 0; id < S; ++id) {
        occ.clear();
        for (int h = 0; h < H; ++h) {
            if (p.paths[h].empty()) continue;
            for (int64_t b = bucket_off[(size_t)id * H + h]; b < bucket_off[(size_t)id * H + h + 1]; ++b) occ.push_back(bucket[(size_t)b]);
        }
        if (occ.size() > 1) std::stable_sort(occ.begin(), occ.end(), [&](int64_t A, int64_t B) { return key_cmp(A, B) < 0; });
        bool all_haps = false;
        for (size_t g0 = 0; g0 < occ.size() && !all_haps;) {
            size_t g1 = g0 + 1;
            while (g1 < occ.size() && key_cmp(occ[g0], occ[g1]) == 0) ++g1;
            if ((float)(g1 - g0) >= threshold * (float)(uint32_t)H) all_haps = true;      // int >= float*uint (:618)
            g0 = g1;
        }
        for (int h = 0; h < H; ++h) kept[h].clear();
        if (!all_haps)
            for (int64_t x : occ) kept[hit_walk[x]].push_back(x);
        for (int h = 0; h < H; ++h) {
            auto& v = kept[h];
            // (first vertex, last vertex), empties last (:643-661); equal keys on one walk are identical lists
            std::stable_sort(v.begin(), v.end(), [&](int64_t A, int64_t B) {
                const bool ea = s.hit_vtx_off[A] == s.hit_vtx_off[A + 1], eb = s.hit_vtx_off[B] == s.hit_vtx_off[B + 1];
                if (ea) return false;
                if (eb) return true;
                const int32_t a0 = s.hit_vtx[s.hit_vtx_off[A]], b0 = s.hit_vtx[s.hit_vtx_off[B]];
                if (a0 != b0) return a0 < b0;
                return s.hit_vtx[s.hit_vtx_off[A + 1] - 1] < s.hit_vtx[s.hit_vtx_off[B + 1] - 1];
            });
            for (int64_t x : v) {
                for (uint64_t t = s.hit_vtx_off[x]; t < s.hit_vtx_off[x + 1]; ++t) a.vtx.push_back(s.hit_vtx[t]);
                a.vtx_off.push_back((int64_t)a.vtx.size());
            }
            a.occ_off[(size_t)id * H + h + 1] = (int64_t)a.vtx_off.size() - 1;
            a.anchors_per_walk[h] += (int64_t)v.size();
        }
    }

    // multiplicity histogram (:745-772) and the classifier (:777-879)
    std::map<int32_t, int32_t> kmer_freq;
    for (int64_t id = 0; id < S; ++id) kmer_freq[(int32_t)s.read_count[id]] += 1;
    std::vector<std::pair<int, double>> hist;
    int max_mult = 0;
    for (const auto& kv : kmer_freq) { hist.emplace_back((int)kv.first, (double)kv.second); max_mult = std::max(max_mult, (int)kv.first); }
    KGFit fit = kg_fit(hist, max_mult, threads);
    char buf[512];
    snprintf(buf, sizeof buf,
             "Fitted model: best NLL=%.2f, u_v=%.2f (hom mean), sd_v=%.2f (hom SD), var_w=%.2f, p_d=%.2f, zp_copy=%.2f, "
             "zp_copy_het=%.2f, err_shape=%.2f, max_copy=%d",
             fit.nll, fit.P.u_v, fit.P.sd_v, fit.P.var_w, fit.P.p_d, fit.P.zp_copy, fit.P.zp_copy_het, fit.P.err_shape, fit.P.max_copy);
    a.fit_line = buf;
    a.homo_bv.assign((size_t)S, 0);
    std::vector<int8_t> label(max_mult + 1, -1);
    for (int64_t id = 0; id < S; ++id) {
        const int m = (int)s.read_count[id];
        if (m == 0) continue;
        if (label[m] < 0) label[m] = kg_is_hom(fit.P, m) ? 1 : 0;
        if (label[m]) { a.homo_bv[id] = 1; ++a.n_hom; } else ++a.n_het;
    }
}

}  // namespace dgh
