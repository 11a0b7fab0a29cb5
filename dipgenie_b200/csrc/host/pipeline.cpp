// pipeline.cpp — see pipeline.h.  Order of the stages = reference src/main.cpp:117-165:
//   gfa_read -> Solver::read_gfa -> read_ip_reads -> compute_and_classify_anchors -> solve
#include "pipeline.h"

#include "../common/dgd_dump.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

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

// Everything before the DP call: parse, panel, device sketch + join, anchors + classifier, expansion, (diploid)
// levelization, flattening into the C-ABI arrays.  `gpu` serialises the device stages of concurrent samples (one
// dg_ctx is used by one host thread at a time).
struct Prepared {
    Options o;
    Panel panel;
    Expanded ex;
    std::vector<int64_t> adj_off, col_off;
    std::vector<int32_t> adj_dst, col_val, level_off;
    std::vector<uint8_t> adj_w;
    RunSummary sum;
    double t0 = 0;
};

// DG_TIMING=1: host stage times on stderr (diagnostics).
struct StageClock {
    const bool on = getenv("DG_TIMING") != nullptr;
    double last = now_s();
    void tick(const char* what) {
        if (!on) return;
        const double t = now_s();
        fprintf(stderr, "[T::host] %-28s %8.1f ms\n", what, (t - last) * 1e3);
        last = t;
    }
};

static int prepare(const Options& o, const Backend& be, Prepared& P, std::mutex* gpu, std::string& err) {
    const double t0 = now_s();
    StageClock clk;
    P.o = o; P.t0 = t0;
    RunSummary& sum = P.sum;
    sum = RunSummary();
    GfaGraph gfa;
    if (!read_gfa_file(o.gfa, gfa, err)) return 1;
    clk.tick("read_gfa_file");
    Panel& panel = P.panel;
    if (!build_panel(gfa, panel, err)) return 1;
    clk.tick("build_panel");
    const int H = (int)panel.paths.size();
    if (o.verbose) fprintf(stderr, "[M::%s::%.3f] loaded the graph: %d segments, %d walks\n", __func__, now_s() - t0, panel.n_vtx, H);
    std::vector<std::string> reads;
    if (!read_sequences(o.reads, reads, err)) return 1;
    clk.tick("read_sequences");

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
        std::unique_lock<std::mutex> lk; if (gpu) lk = std::unique_lock<std::mutex>(*gpu);
        const int e = be.sketch_reads(be.ctx, (const uint8_t*)bases.data(), off.data(), (uint32_t)reads.size(), o.k, o.w, &sp.p, &rc.p, &ns);
        if (e) { err = std::string("dg_sketch_reads: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        sk.spectrum.assign(sp.p, sp.p + ns);
        sk.read_count.assign(rc.p, rc.p + ns);
    }
    clk.tick("stage: sketch_reads");
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
        std::unique_lock<std::mutex> lk; if (gpu) lk = std::unique_lock<std::mutex>(*gpu);
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

    clk.tick("stage: index_walks");
    // ---- host: filter, occurrence order, hom/het classifier (solver.cpp:590-887) ----
    Anchors anchors;
    build_anchors(panel, sk, o.threshold, o.threads, anchors);
    clk.tick("build_anchors (+ fit)");
    sum.n_hom = anchors.n_hom; sum.n_het = anchors.n_het;
    if (o.verbose) {
        for (int h = 0; h < H; ++h)
            fprintf(stderr, "Haplotype: %s Number of Anchors: %lld\n", panel.walk_names[h].c_str(), (long long)anchors.anchors_per_walk[h]);
        fprintf(stderr, "[M::%s::%.3f] %s\n", __func__, now_s() - t0, anchors.fit_line.c_str());
    }

    // ---- host: haplotype-expanded graph in Kahn order (approximator.cpp:1014-1256) ----
    Expanded& ex = P.ex;
    if (!expand_graph(panel, anchors, ex, err)) return 1;
    clk.tick("expand_graph");
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

    if (o.ploidy != 1) {
        if (!levelize(ex.g, err)) return 1;      // ExpandedGraph.hpp:269-409
        const int L = (int)ex.g.vertices_in_level.size();
        P.level_off.assign((size_t)L + 1, 0);
        for (int l = 0; l < L; ++l) P.level_off[(size_t)l + 1] = P.level_off[l] + (int32_t)ex.g.vertices_in_level[l].size();
    }
    clk.tick("levelize");
    flatten(ex.g, P.adj_off, P.adj_dst, P.adj_w, P.col_off, P.col_val);
    clk.tick("flatten");
    return 0;
}

// After the diploid DP: stitch the two sequences and write the FASTA (approximator.cpp:779-925, :1314-1325).
static int finish_diploid(Prepared& P, int32_t value, const int32_t* e1, int32_t n1, const int32_t* e2, int32_t n2, std::string& err) {
    const Options& o = P.o;
    P.sum.dp_value = value;
    if (o.verbose) fprintf(stderr, "DP value: %d\n", value);
    std::vector<std::pair<int, int>> p1, p2;
    for (int x = 0; x < n1; ++x) p1.emplace_back(e1[2 * x], e1[2 * x + 1]);
    for (int x = 0; x < n2; ++x) p2.emplace_back(e2[2 * x], e2[2 * x + 1]);
    DiploidSolution sol;
    if (!stitch_diploid(P.panel, P.ex.g, p1, p2, sol, err)) return 1;
    P.sum.r1 = sol.r1; P.sum.r2 = sol.r2; P.sum.len1 = (int64_t)sol.hap1.size(); P.sum.len2 = (int64_t)sol.hap2.size();
    if (o.verbose)
        fprintf(stderr, "Recombinations in P1: %d, P2: %d, bp: %lld / %lld\n", sol.r1, sol.r2, (long long)P.sum.len1, (long long)P.sum.len2);
    if (!write_fasta_diploid(o.out, sol.hap1, sol.hap2)) { err = "cannot write " + o.out; return 1; }
    return 0;
}

static int solve_haploid(Prepared& P, const Backend& be, std::string& err) {
    const Options& o = P.o;
    // ---- device stage 3a: haploid DP + R+1 tracebacks (approximator.cpp:44-102, :141-153) ----
    std::vector<int32_t> colours((size_t)o.R + 1, 0);
    std::vector<int64_t> path_off((size_t)o.R + 2, 0);
    Owned<int32_t> paths;
    paths.fr = be.free_array;
    const int e = be.dp_haploid(be.ctx, (int32_t)P.ex.g.adj.size(), P.adj_off.data(), P.adj_dst.data(), P.adj_w.data(), P.col_off.data(),
                                P.col_val.data(), P.ex.n_colours, o.R, colours.data(), path_off.data(), &paths.p);
    if (e) { err = std::string("dg_dp_haploid: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
    std::string log;
    const int best_r = haploid_best_r(std::vector<int>(colours.begin(), colours.end()), log);   // double arithmetic stays on the host
    P.sum.best_r = best_r;
    if (o.verbose) fprintf(stderr, "%sRecombination count: %d\n", log.c_str(), best_r);
    std::vector<int32_t> path(paths.p + path_off[(size_t)best_r], paths.p + path_off[(size_t)best_r + 1]);
    const std::string seq = haploid_sequence(P.panel, P.ex.g, path);
    P.sum.len1 = (int64_t)seq.size();
    if (!write_fasta_haploid(o.out, seq)) { err = "cannot write " + o.out; return 1; }
    return 0;
}

int run_pipeline(const Options& o, const Backend& be, RunSummary& sum, std::string& err) {
    Prepared P;
    if (int rc = prepare(o, be, P, nullptr, err)) return rc;
    StageClock clk;
    int rc = 0;
    if (o.ploidy == 1) {
        rc = solve_haploid(P, be, err);
    } else {
        // Tooling (bench.py's wide-panel workloads, tests): DG_DUMP_DIPIN=<file> writes the DP's input — the levelized graph
        // this front end built — as a DGD1 container; DG_DUMP_ONLY=1 stops there.
        if (const char* dump = getenv("DG_DUMP_DIPIN")) {
            dgd::Writer w(dump);
            if (!w.ok()) { err = std::string("cannot write ") + dump; return 1; }
            w.put("level_off", P.level_off); w.put("adj_off", P.adj_off); w.put("adj_dst", P.adj_dst); w.put("adj_w", P.adj_w);
            w.put("col_off", P.col_off); w.put("col_val", P.col_val);
            w.put("colour_is_hom", P.ex.color_homo_bv);
            if (getenv("DG_DUMP_ONLY")) { sum = P.sum; return 0; }
        }
        // ---- device stage 3b: diploid DP + edge lists ----
        int32_t value = 0, s_het = 0, n1 = 0, n2 = 0;
        std::vector<int32_t> e1(2 * ((size_t)o.R + 2)), e2(2 * ((size_t)o.R + 2));
        const int e = be.dp_diploid(be.ctx, (int32_t)P.level_off.size() - 1, P.level_off.data(), P.adj_off.data(), P.adj_dst.data(),
                                    P.adj_w.data(), P.col_off.data(), P.col_val.data(), P.ex.color_homo_bv.data(),
                                    (int32_t)P.ex.color_homo_bv.size(), o.R, &value, &s_het, e1.data(), &n1, e2.data(), &n2);
        if (e) { err = std::string("dg_dp_diploid: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); return 1; }
        clk.tick("stage: dp_diploid");
        rc = finish_diploid(P, value, e1.data(), n1, e2.data(), n2, err);
        clk.tick("stitch + FASTA");
    }
    sum = P.sum;
    if (!rc && o.verbose) fprintf(stderr, "[M::%s] Real time: %.3f sec\n", __func__, now_s() - P.t0);
    return rc;
}

// Many samples in one process (the reference's batch script starts one process per sample,
// data/run_DipGenie_batch.sh:21-39): the stages before the DP run sample-parallel on the host threads (device
// sketch stages one at a time), all diploid DPs go to the GPU together (dg_dp_diploid_batch), stitching and FASTA
// output run sample-parallel again.
int run_batch(const std::vector<Options>& jobs, const Backend& be, std::vector<RunSummary>& sums, std::vector<std::string>& errs) {
    const int n = (int)jobs.size();
    sums.assign((size_t)n, RunSummary());
    errs.assign((size_t)n, std::string());
    std::vector<std::unique_ptr<Prepared>> P((size_t)n);
    std::vector<int> rc((size_t)n, 0);
    std::mutex gpu;
    const int nt = std::max(1, std::min(n, jobs.empty() ? 1 : jobs[0].threads));
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int i = 0; i < n; ++i) {
        P[(size_t)i].reset(new Prepared());
        rc[(size_t)i] = prepare(jobs[(size_t)i], be, *P[(size_t)i], &gpu, errs[(size_t)i]);
    }
    std::vector<int> dip;
    for (int i = 0; i < n; ++i) {
        if (rc[(size_t)i]) continue;
        if (jobs[(size_t)i].ploidy == 1) rc[(size_t)i] = solve_haploid(*P[(size_t)i], be, errs[(size_t)i]);
        else dip.push_back(i);
    }
    std::vector<dg_dip_output_t> out(dip.size());
    if (!dip.empty() && be.dp_diploid_batch) {
        std::vector<dg_dip_input_t> in(dip.size());
        for (size_t x = 0; x < dip.size(); ++x) {
            Prepared& q = *P[(size_t)dip[x]];
            in[x] = {(int32_t)q.level_off.size() - 1, q.level_off.data(), q.adj_off.data(), q.adj_dst.data(), q.adj_w.data(), q.col_off.data(),
                     q.col_val.data(), q.ex.color_homo_bv.data(), (int32_t)q.ex.color_homo_bv.size(), q.o.R};
        }
        be.dp_diploid_batch(be.ctx, (int32_t)dip.size(), in.data(), out.data(), 0, 0);
        for (size_t x = 0; x < dip.size(); ++x)
            if (out[x].status) { rc[(size_t)dip[x]] = 1; errs[(size_t)dip[x]] = std::string("dg_dp_diploid_batch: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); }
    } else {
        for (size_t x = 0; x < dip.size(); ++x) {
            Prepared& q = *P[(size_t)dip[x]];
            dg_dip_output_t& o = out[x];
            memset(&o, 0, sizeof o);
            if (q.o.R + 2 > DG_BATCH_MAX_EDGES) { rc[(size_t)dip[x]] = 1; errs[(size_t)dip[x]] = "R too large for the batch record"; continue; }
            o.status = be.dp_diploid(be.ctx, (int32_t)q.level_off.size() - 1, q.level_off.data(), q.adj_off.data(), q.adj_dst.data(), q.adj_w.data(),
                                     q.col_off.data(), q.col_val.data(), q.ex.color_homo_bv.data(), (int32_t)q.ex.color_homo_bv.size(), q.o.R,
                                     &o.sink_value, &o.sink_s_het, o.p1_edges, &o.n_p1, o.p2_edges, &o.n_p2);
            if (o.status) { rc[(size_t)dip[x]] = 1; errs[(size_t)dip[x]] = std::string("dg_dp_diploid: ") + (be.last_error ? be.last_error(be.ctx) : "failed"); }
        }
    }
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int x = 0; x < (int)dip.size(); ++x) {
        const int i = dip[(size_t)x];
        if (rc[(size_t)i]) continue;
        const dg_dip_output_t& o = out[(size_t)x];
        rc[(size_t)i] = finish_diploid(*P[(size_t)i], o.sink_value, o.p1_edges, o.n_p1, o.p2_edges, o.n_p2, errs[(size_t)i]);
    }
    int bad = 0;
    for (int i = 0; i < n; ++i) { sums[(size_t)i] = P[(size_t)i] ? P[(size_t)i]->sum : RunSummary(); if (rc[(size_t)i]) ++bad; }
    return bad;
}

}  // namespace dgh
