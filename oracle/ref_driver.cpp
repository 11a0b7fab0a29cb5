// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A driver written for this repo that links the *reference's own* object files
// (built by oracle/Makefile from the sources where they lie under
// /root/reference/src; approximator.cpp from an instrumented scratch copy that
// only adds the read-only dump hooks of ref_hook.h) and walks the reference's
// public seams exactly the way its main() does (src/main.cpp:117-165):
//   gfa_read -> Approximator::read_gfa -> read_ip_reads ->
//   compute_and_classify_anchors -> solve
// while writing every intermediate product of the hot path to a DGD1 file:
//   panel model (src/solver.cpp:27-227), per-walk minimizer index
//   (Solver::index_kmers, :277-363), per-read hash sets (Solver::compute_hashes,
//   :366-412), filtered Anchor_hits + homo_bv (:449-887), the ExpandedGraph that
//   enters the DP, per-level DP checksums, recombination-edge lists, and results.
//
// Usage: ref_driver -g graph.gfa[.gz] -r reads.fq[.gz] -o out.fa -D dump.dgd
//                   [-p 1|2] [-R 18] [-k 31] [-w 25] [-t 4] [-T 1.0] [-S (skip solve)]
//                   [-L (light: skip per-walk index / per-read hashes in the dump)]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unistd.h>

#include "gfa-priv.h"
#include "PHIpriv.h"
#include "approximator.h"
#include "misc.h"
#include "sys.h"

#include "ref_hook.h"
#include "../dipgenie_b200/csrc/common/dgd_dump.h"

static std::unique_ptr<dgd::Writer> g_w;
static std::vector<uint64_t> g_lvl_sum, g_lvl_live;

static void dump_graph(const char* prefix, const ExpandedGraph& g) {
    if (!g_w) return;
    std::string p(prefix);
    std::vector<int64_t> off; off.push_back(0);
    std::vector<int32_t> dst, wt;
    for (const auto& nb : g.adj_list) {
        for (const auto& e : nb) { dst.push_back(e.first); wt.push_back(e.second); }
        off.push_back((int64_t)dst.size());
    }
    g_w->put(p + "adj.off", off);
    g_w->put(p + "adj.dst", dst);
    g_w->put(p + "adj.w", wt);
    g_w->put_ragged_i32(p + "color", g.color);
    g_w->put_ragged_i32(p + "orig", g.original_vertex);
    std::vector<int32_t> hap(g.haplotype.begin(), g.haplotype.end());
    g_w->put(p + "hap", hap);
    std::vector<int32_t> lvl(g.level.begin(), g.level.end());
    g_w->put(p + "level", lvl);
    g_w->put_ragged_i32(p + "lvl_vtx", g.vertices_in_level);
}

void dg_ref_dump_haploid_input(const ExpandedGraph& g, int R) {
    if (!g_w) return;
    dump_graph("hap_in.", g);
    g_w->put_i64("hap_in.R", R);
}

void dg_ref_dump_diploid_input(const ExpandedGraph& g, int R, const std::vector<bool>& color_homo_bv) {
    if (!g_w) return;
    dump_graph("dip_in.", g);
    g_w->put_i64("dip_in.R", R);
    std::vector<uint8_t> bv(color_homo_bv.size());
    for (size_t i = 0; i < bv.size(); ++i) bv[i] = color_homo_bv[i] ? 1 : 0;
    g_w->put("dip_in.color_homo", bv);
}

void dg_ref_dump_haploid_result(const std::vector<int>& colors_by_r, int best_r,
                                const std::vector<int>& path, const std::vector<int>& path_original) {
    if (!g_w) return;
    g_w->put("hap_out.colors_by_r", std::vector<int32_t>(colors_by_r.begin(), colors_by_r.end()));
    g_w->put_i64("hap_out.best_r", best_r);
    g_w->put("hap_out.path", std::vector<int32_t>(path.begin(), path.end()));
    g_w->put("hap_out.path_original", std::vector<int32_t>(path_original.begin(), path_original.end()));
}

void dg_ref_dump_diploid_result(const std::vector<std::pair<int, int>>& p1,
                                const std::vector<std::pair<int, int>>& p2, int value, int s_het) {
    if (!g_w) return;
    std::vector<int32_t> a, b;
    for (auto& e : p1) { a.push_back(e.first); a.push_back(e.second); }
    for (auto& e : p2) { b.push_back(e.first); b.push_back(e.second); }
    g_w->put("dip_out.p1_edges", a);
    g_w->put("dip_out.p2_edges", b);
    g_w->put_i64("dip_out.value", value);
    g_w->put_i64("dip_out.s_het", s_het);
    g_w->put("dip_out.level_checksum", g_lvl_sum);
    g_w->put("dip_out.level_live", g_lvl_live);
}

void dg_ref_level_checksum_push(int level, uint64_t checksum, uint64_t n_live) {
    if ((size_t)level >= g_lvl_sum.size()) { g_lvl_sum.resize(level + 1, 0); g_lvl_live.resize(level + 1, 0); }
    g_lvl_sum[level] = checksum;
    g_lvl_live[level] = n_live;
}

// DP-only mode (-G graph.dgd): times the reference's own diploid solver
// (Approximator::diploid_dp_approximation_solver, src/approximator.cpp:362) on a levelized graph given
// as flat arrays (LevelGraph layout, see include/dipgenie_cuda.h), optionally truncated to its first
// -M levels (a bounded sample of the same workload: a sink is appended after level M-1).
// The sequence-stitching inputs are stubbed (one empty segment), so only the sweep + list recovery run.
static int dp_only(const std::string& path, int R, int threads, int max_levels) {
    dgd::Reader rd(path);
    if (!rd.ok()) { fprintf(stderr, "ref_driver: cannot read %s\n", path.c_str()); return 1; }
    std::vector<int32_t> level_off, adj_dst, col_val;
    std::vector<int64_t> adj_off, col_off;
    std::vector<uint8_t> adj_w, hom;
    if (!rd.get("level_off", level_off) || !rd.get("adj_off", adj_off) || !rd.get("adj_dst", adj_dst) ||
        !rd.get("adj_w", adj_w) || !rd.get("col_off", col_off) || !rd.get("col_val", col_val) ||
        !rd.get("colour_is_hom", hom)) { fprintf(stderr, "ref_driver: graph arrays missing\n"); return 1; }
    int L = (int)level_off.size() - 1;
    int keep = (max_levels > 1 && max_levels < L) ? max_levels : L;
    const int32_t V = level_off[keep];
    const bool cut = keep < L;
    ExpandedGraph g;
    const int32_t n = V + (cut ? 1 : 0);
    g.adj_list.resize(n); g.color.resize(n); g.original_vertex.assign(n, std::vector<int>{0});
    g.haplotype.assign(n, 0); g.level.assign(n, 0); g.vertices_in_level.resize(keep + (cut ? 1 : 0));
    unsigned long long U = 0;
    for (int l = 0; l < keep; ++l) {
        unsigned long long El = 0;
        for (int32_t v = level_off[l]; v < level_off[l + 1]; ++v) {
            g.level[v] = l; g.vertices_in_level[l].push_back(v);
            for (int64_t c = col_off[v]; c < col_off[v + 1]; ++c) g.color[v].push_back(col_val[c]);
            if (cut && l == keep - 1) { g.adj_list[v].push_back({V, 0}); El += 1; continue; }
            for (int64_t e = adj_off[v]; e < adj_off[v + 1]; ++e) g.adj_list[v].push_back({adj_dst[e], (int)adj_w[e]});
            El += (unsigned long long)(adj_off[v + 1] - adj_off[v]);
        }
        U += (unsigned long long)(R + 1) * El * El;
    }
    if (cut) { g.level[V] = keep; g.vertices_in_level[keep].push_back(V); }
    std::vector<bool> bv(hom.size());
    for (size_t i = 0; i < hom.size(); ++i) bv[i] = hom[i] != 0;
    Approximator* A = new Approximator(nullptr);
    A->num_threads = threads; A->recombination_limit = R;
    A->paths.assign(1, std::vector<uint32_t>{0}); A->node_seq.assign(1, std::string());
    std::vector<std::vector<Approximator::AnchorRec>> anchorsByHap(1);
    double t0 = realtime();
    auto sol = A->diploid_dp_approximation_solver(g, R, bv, anchorsByHap);
    double ms = (realtime() - t0) * 1e3;
    int r1 = sol.empty() ? -1 : std::get<0>(sol[0]), r2 = sol.empty() ? -1 : std::get<1>(sol[0]);
    printf("\nDPONLY levels %d threads %d R %d cell_updates %llu ms %.3f r1 %d r2 %d\n", keep + (cut ? 1 : 0), threads, R, U, ms, r1, r2);
    return 0;
}

int main(int argc, char** argv) {
    std::string gfa, reads, out = "/dev/null", dump, graph;
    int ploidy = 2, R = 18, k = 31, w = 25, threads = 4, skip_solve = 0, light = 0, max_levels = 0;
    float threshold = 1.0f;
    int c;
    while ((c = getopt(argc, argv, "g:r:o:D:p:R:k:w:t:T:SLG:M:")) >= 0) {
        switch (c) {
            case 'G': graph = optarg; break;
            case 'M': max_levels = atoi(optarg); break;
            case 'g': gfa = optarg; break;
            case 'r': reads = optarg; break;
            case 'o': out = optarg; break;
            case 'D': dump = optarg; break;
            case 'p': ploidy = atoi(optarg); break;
            case 'R': R = atoi(optarg); break;
            case 'k': k = atoi(optarg); break;
            case 'w': w = atoi(optarg); break;
            case 't': threads = atoi(optarg); break;
            case 'T': threshold = (float)atof(optarg); break;
            case 'S': skip_solve = 1; break;
            case 'L': light = 1; break;
            default: return 2;
        }
    }
    if (!graph.empty()) return dp_only(graph, R, threads, max_levels);
    if (gfa.empty() || reads.empty()) { fprintf(stderr, "ref_driver: -g and -r required\n"); return 2; }
    if (!dump.empty()) g_w.reset(new dgd::Writer(dump));

    mg_realtime0 = realtime();
    gfa_t* g = gfa_read(gfa.c_str());
    if (!g) { fprintf(stderr, "ref_driver: cannot read %s\n", gfa.c_str()); return 1; }

    Approximator* A = new Approximator(g);
    A->read_gfa();
    A->num_threads = threads;
    A->hap_file = out;
    A->debug = false;
    A->hap_name = "ref_driver";
    A->k_mer = k;
    A->window = w;
    A->bucket_bits = 14;
    A->max_occ = 5000;
    A->recombination_limit = R;
    A->recombination_penalty = 100;
    A->is_qclp = 1;
    A->is_naive_exp = 0;
    A->threshold = threshold;
    A->is_mixed = true;

    std::vector<std::pair<std::string, std::string>> ip_reads;
    A->read_ip_reads(ip_reads, reads);

    if (g_w) {
        // panel model: src/solver.cpp:27-227
        g_w->put_i64("panel.n_vtx", A->n_vtx);
        g_w->put_i64("panel.num_walks", A->num_walks);
        g_w->put_ragged_i32("panel.paths", A->paths);
        g_w->put_ragged_i32("panel.adj", A->adj_list);
        g_w->put("panel.top_order_map", std::vector<int32_t>(A->top_order_map.begin(), A->top_order_map.end()));
        std::vector<int32_t> nl(A->node_seq.size());
        for (size_t i = 0; i < nl.size(); ++i) nl[i] = (int32_t)A->node_seq[i].size();
        g_w->put("panel.node_len", nl);
        std::string names;
        for (auto& s : A->hap_id2name) { names += s; names += '\n'; }
        g_w->put_str("panel.walk_names", names);
        g_w->put_i64("reads.n", (int64_t)ip_reads.size());
        if (!light) {
            // raw inputs of the sketch stage (node_seq, src/solver.cpp:63-70; read sequences, :230-245) so
            // that the sketch fixtures are self-contained
            std::string segs, rds;
            std::vector<int64_t> so, ro; so.push_back(0); ro.push_back(0);
            for (auto& q : A->node_seq) { segs += q; so.push_back((int64_t)segs.size()); }
            for (auto& rd : ip_reads) { rds += rd.second; ro.push_back((int64_t)rds.size()); }
            g_w->put_str("panel.node_seq", segs);
            g_w->put("panel.node_seq_off", so);
            g_w->put_str("reads.bases", rds);
            g_w->put("reads.off", ro);
        }

        if (!light) {
            // per-walk minimizer index: Solver::index_kmers, src/solver.cpp:277-363
            for (uint32_t h = 0; h < A->num_walks; ++h) {
                auto idx = A->index_kmers((int32_t)h);
                std::vector<uint64_t> hs; hs.reserve(idx.size());
                std::vector<std::vector<int32_t>> vt; vt.reserve(idx.size());
                for (auto& e : idx) { hs.push_back(e.first); vt.push_back(e.second.k_mers); }
                g_w->put("index." + std::to_string(h) + ".hash", hs);
                g_w->put_ragged_i32("index." + std::to_string(h) + ".vtx", vt);
            }
            // per-read hash sets: Solver::compute_hashes, src/solver.cpp:366-412 (works on a copy;
            // the reference mutates its argument to upper case)
            std::vector<int64_t> off; off.push_back(0);
            std::vector<uint64_t> val;
            for (auto& rd : ip_reads) {
                std::string s = rd.second;
                auto hs = A->compute_hashes(s);
                for (auto x : hs) val.push_back(x);
                off.push_back((int64_t)val.size());
            }
            g_w->put("read_hashes.off", off);
            g_w->put("read_hashes.val", val);
        }
    }

    A->compute_and_classify_anchors(ip_reads);

    if (g_w) {
        g_w->put_i64("anchors.count_sp_r", A->count_sp_r);
        // Anchor_hits[id][h] -> list of vertex lists (after filter + sort, solver.cpp:590-663)
        std::vector<int64_t> occ_off; occ_off.push_back(0);   // per (id,h): range into occ list
        std::vector<int64_t> vtx_off; vtx_off.push_back(0);   // per occ: range into vtx
        std::vector<int32_t> vtx;
        for (size_t a = 0; a < A->Anchor_hits.size(); ++a)
            for (size_t h = 0; h < A->Anchor_hits[a].size(); ++h) {
                for (auto& occ : A->Anchor_hits[a][h]) {
                    for (auto v : occ) vtx.push_back(v);
                    vtx_off.push_back((int64_t)vtx.size());
                }
                occ_off.push_back((int64_t)vtx_off.size() - 1);
            }
        g_w->put("anchors.occ_off", occ_off);
        g_w->put("anchors.vtx_off", vtx_off);
        g_w->put("anchors.vtx", vtx);
        std::vector<uint8_t> bv(A->homo_bv.size());
        for (size_t i = 0; i < bv.size(); ++i) bv[i] = A->homo_bv[i] ? 1 : 0;
        g_w->put("anchors.homo_bv", bv);
    }

    if (!skip_solve) {
        if (ploidy != 1 && ploidy != 2) { fprintf(stderr, "ref_driver: ploidy must be 1 or 2\n"); return 0; }
        A->solve(ip_reads, ploidy == 2);
    }
    g_w.reset();
    return 0;
}
