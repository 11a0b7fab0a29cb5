// pipeline.cpp — see pipeline.h.  Order of the stages = reference src/main.cpp:117-165:
//   gfa_read -> Solver::read_gfa -> read_ip_reads -> compute_and_classify_anchors -> solve
#include "pipeline.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dgh.h"

namespace dgh {

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class T>
struct Owned {                         // library-allocated array, released with the backend's free
    T* p = nullptr;
    void (*fr)(void*) = nullptr;
    ~Owned() { if (p && fr) fr(p); }
};
}  // namespace

int run_pipeline(const Options& o, const Backend& be, RunSummary& sum, std::string& err) {
    const double t0 = now_s();
    sum = RunSummary();
    GfaGraph gfa;
    if (!read_gfa_file(o.gfa, gfa, err)) return 1;
    Panel panel;
    if (!build_panel(gfa, panel, err)) return 1;
    const int H = (int)panel.paths.size();
    if (o.verbose) fprintf(stderr, "[M::%s::%.3f] loaded the graph: %d segments, %d walks\n", __func__, now_s() - t0, panel.n_vtx, H);
    std::vector<std::string> reads;
    if (!read_sequences(o.reads, reads, err)) return 1;

    // ---- device stage 1: read sketch -> spectrum + multiplicities (solver.cpp:526-558, :711-732) ----
    SketchResult sk;
    {
        std::vector<uint64_t> off(reads.size() + 1, 0);
        for (size_t i = 0; i < reads.size(); ++i) off[i + 1] = off[i] + reads[i].size();
        std::string bases;
        bases.reserve((size_t)off.back());
        for (auto& r : reads) bases += r;
        Owned<uint64_t> sp; Owned<uint32_t> rc;
        sp.fr = rc.fr = be.free_array;
        uint64_t ns = 0;
        const int e = be.sketch_reads(be.ctx, (const uint8_t*)bases.data(), off.data(), (uint32_t)reads.size(), o.k, o.w, &sp.p, &rc.p, &ns);
        if (e) { err = std::string("dg_sketch_reads: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        sk.spectrum.assign(sp.p, sp.p + ns);
        sk.read_count.assign(rc.p, rc.p + ns);
    }
    sum.spectrum = (int64_t)sk.spectrum.size();
    if (o.verbose) fprintf(stderr, "[M::%s::%.3f] Count_Sp_R : %lld\n", __func__, now_s() - t0, (long long)sum.spectrum);

    // ---- device stage 2: walk index joined with the spectrum (solver.cpp:277-363, :415-446, :560-576) ----
    {
        std::vector<uint64_t> seg_off((size_t)panel.n_vtx + 1, 0);
        for (int v = 0; v < panel.n_vtx; ++v) seg_off[(size_t)v + 1] = seg_off[v] + panel.node_seq[v].size();
        std::string seg_bases;
        seg_bases.reserve((size_t)seg_off.back());
        for (auto& s : panel.node_seq) seg_bases += s;
        std::vector<uint64_t> walk_off((size_t)H + 1, 0);
        std::vector<int32_t> walk_vtx;
        for (int h = 0; h < H; ++h) { walk_vtx.insert(walk_vtx.end(), panel.paths[h].begin(), panel.paths[h].end()); walk_off[(size_t)h + 1] = walk_vtx.size(); }
        sk.n_minimizers.assign((size_t)std::max(H, 1), 0);
        Owned<uint64_t> ho, vo; Owned<uint32_t> hs; Owned<int32_t> hv;
        ho.fr = vo.fr = hs.fr = hv.fr = be.free_array;
        const int e = be.index_walks(be.ctx, (const uint8_t*)seg_bases.data(), seg_off.data(), (uint32_t)panel.n_vtx, walk_vtx.data(),
                                     walk_off.data(), (uint32_t)H, panel.top_order_map.data(), o.k, o.w, sk.spectrum.data(),
                                     (uint64_t)sk.spectrum.size(), sk.n_minimizers.data(), &ho.p, &hs.p, &vo.p, &hv.p);
        if (e) { err = std::string("dg_index_walks: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        sk.hit_off.assign(ho.p, ho.p + H + 1);
        const uint64_t nh = sk.hit_off[(size_t)H];
        sk.hit_sid.assign(hs.p, hs.p + nh);
        sk.hit_vtx_off.assign(vo.p, vo.p + nh + 1);
        sk.hit_vtx.assign(hv.p, hv.p + sk.hit_vtx_off[nh]);
        if (o.verbose)
            for (int h = 0; h < H; ++h)
                fprintf(stderr, "Haplotype: %s Number of Minimizers: %llu\n", panel.walk_names[h].c_str(), (unsigned long long)sk.n_minimizers[h]);
    }

    // ---- host: filter, occurrence order, hom/het classifier (solver.cpp:590-887) ----
    Anchors anchors;
    build_anchors(panel, sk, o.threshold, o.threads, anchors);
    sum.n_hom = anchors.n_hom; sum.n_het = anchors.n_het;
    if (o.verbose) {
        for (int h = 0; h < H; ++h)
            fprintf(stderr, "Haplotype: %s Number of Anchors: %lld\n", panel.walk_names[h].c_str(), (long long)anchors.anchors_per_walk[h]);
        fprintf(stderr, "[M::%s::%.3f] %s\n", __func__, now_s() - t0, anchors.fit_line.c_str());
    }

    // ---- host: haplotype-expanded graph in Kahn order (approximator.cpp:1014-1256) ----
    Expanded ex;
    if (!expand_graph(panel, anchors, ex, err)) return 1;
    auto flatten = [](const ExpGraph& g, std::vector<int64_t>& adj_off, std::vector<int32_t>& adj_dst, std::vector<uint8_t>& adj_w,
                      std::vector<int64_t>& col_off, std::vector<int32_t>& col_val) {
        const size_t n = g.adj.size();
        adj_off.assign(n + 1, 0); col_off.assign(n + 1, 0);
        adj_dst.clear(); adj_w.clear(); col_val.clear();
        for (size_t u = 0; u < n; ++u) {
            for (auto& e : g.adj[u]) { adj_dst.push_back(e.first); adj_w.push_back((uint8_t)e.second); }
            adj_off[u + 1] = (int64_t)adj_dst.size();
            col_val.insert(col_val.end(), g.color[u].begin(), g.color[u].end());
            col_off[u + 1] = (int64_t)col_val.size();
        }
    };
    std::vector<int64_t> adj_off, col_off;
    std::vector<int32_t> adj_dst, col_val;
    std::vector<uint8_t> adj_w;

    if (o.ploidy == 1) {
        // ---- device stage 3a: haploid DP + R+1 tracebacks (approximator.cpp:44-102, :141-153) ----
        flatten(ex.g, adj_off, adj_dst, adj_w, col_off, col_val);
        std::vector<int32_t> colours((size_t)o.R + 1, 0);
        std::vector<int64_t> path_off((size_t)o.R + 2, 0);
        Owned<int32_t> paths;
        paths.fr = be.free_array;
        const int e = be.dp_haploid(be.ctx, (int32_t)ex.g.adj.size(), adj_off.data(), adj_dst.data(), adj_w.data(), col_off.data(),
                                    col_val.data(), ex.n_colours, o.R, colours.data(), path_off.data(), &paths.p);
        if (e) { err = std::string("dg_dp_haploid: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        std::string log;
        const int best_r = haploid_best_r(std::vector<int>(colours.begin(), colours.end()), log);   // double arithmetic stays on the host
        sum.best_r = best_r;
        if (o.verbose) fprintf(stderr, "%sRecombination count: %d\n", log.c_str(), best_r);
        std::vector<int32_t> path(paths.p + path_off[(size_t)best_r], paths.p + path_off[(size_t)best_r + 1]);
        const std::string seq = haploid_sequence(panel, ex.g, path);
        sum.len1 = (int64_t)seq.size();
        if (!write_fasta_haploid(o.out, seq)) { err = "cannot write " + o.out; return 1; }
    } else {
        // ---- host: levelization (ExpandedGraph.hpp:269-409); device stage 3b: diploid DP + edge lists ----
        if (!levelize(ex.g, err)) return 1;
        flatten(ex.g, adj_off, adj_dst, adj_w, col_off, col_val);
        const int L = (int)ex.g.vertices_in_level.size();
        std::vector<int32_t> level_off((size_t)L + 1, 0);
        for (int l = 0; l < L; ++l) level_off[(size_t)l + 1] = level_off[l] + (int32_t)ex.g.vertices_in_level[l].size();
        int32_t value = 0, s_het = 0, n1 = 0, n2 = 0;
        std::vector<int32_t> e1(2 * ((size_t)o.R + 2)), e2(2 * ((size_t)o.R + 2));
        const int e = be.dp_diploid(be.ctx, L, level_off.data(), adj_off.data(), adj_dst.data(), adj_w.data(), col_off.data(), col_val.data(),
                                    ex.color_homo_bv.data(), (int32_t)ex.color_homo_bv.size(), o.R, &value, &s_het, e1.data(), &n1,
                                    e2.data(), &n2);
        if (e) { err = std::string("dg_dp_diploid: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        sum.dp_value = value;
        if (o.verbose) fprintf(stderr, "DP value: %d\n", value);
        std::vector<std::pair<int, int>> p1, p2;
        for (int x = 0; x < n1; ++x) p1.emplace_back(e1[2 * x], e1[2 * x + 1]);
        for (int x = 0; x < n2; ++x) p2.emplace_back(e2[2 * x], e2[2 * x + 1]);
        DiploidSolution sol;
        if (!stitch_diploid(panel, ex.g, p1, p2, sol, err)) return 1;
        sum.r1 = sol.r1; sum.r2 = sol.r2; sum.len1 = (int64_t)sol.hap1.size(); sum.len2 = (int64_t)sol.hap2.size();
        if (o.verbose)
            fprintf(stderr, "Recombinations in P1: %d, P2: %d, bp: %lld / %lld\n", sol.r1, sol.r2, (long long)sum.len1, (long long)sum.len2);
        if (!write_fasta_diploid(o.out, sol.hap1, sol.hap2)) { err = "cannot write " + o.out; return 1; }
    }
    if (o.verbose) fprintf(stderr, "[M::%s] Real time: %.3f sec\n", __func__, now_s() - t0);
    return 0;
}

}  // namespace dgh
