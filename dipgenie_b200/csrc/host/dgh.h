// dgh.h — internal C++ interfaces of the host glue (libdipgenie_host.so).
//
// The host keeps what the reference keeps on the host (north_star: "the C++ host ... stays the drop-in
// front end and calls CUDA through a thin C-ABI layer"): GFA / read parsing, the panel model, the
// anchor filter and hom/het classifier, the haplotype-expanded graph, Kahn order, levelization, sequence
// stitching and FASTA output.  Everything here is written for this repo; each function cites the
// reference lines whose behaviour it reproduces (bit-exact outputs are the contract).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace dgh {

// ---------------------------------------------------------------------------------- GFA (gfa.cpp)
struct GfaWalk {
    std::string sample;
    int hap = 0;
    std::vector<uint32_t> v;   // oriented vertices: segment << 1 | reverse
};
struct GfaGraph {
    std::vector<std::string> seg_name;
    std::vector<std::string> seg_seq;
    std::vector<int32_t> seg_len;
    std::vector<uint8_t> seg_del;
    std::vector<std::pair<uint32_t, uint32_t>> arcs;   // oriented (v, w), complements included, deleted removed
    std::vector<GfaWalk> walks;
};
// gfa_read (reference src/gfa-io.cpp:462-508 with gfa_finalize, src/gfa-base.cpp:421-430)
bool read_gfa_file(const std::string& path, GfaGraph& g, std::string& err);

// ---------------------------------------------------------------------------------- reads (reads.cpp)
// Solver::read_ip_reads (src/solver.cpp:230-245): FASTA/FASTQ, optionally gzip, whole file in memory.
bool read_sequences(const std::string& path, std::vector<std::string>& seqs, std::string& err);

// ---------------------------------------------------------------------------------- panel (panel.cpp)
struct Panel {                       // products of Solver::read_gfa (src/solver.cpp:27-227)
    int32_t n_vtx = 0;               // forward-strand vertices (= segments)
    std::vector<std::string> node_seq;
    std::vector<std::vector<int32_t>> adj;        // sorted by (dense column, id) (:216-223)
    std::vector<std::vector<int32_t>> paths;      // walk -> vertex ids
    std::vector<std::string> walk_names;          // sample.hap (:110)
    std::vector<int32_t> top_order_map;           // vertex -> rank (:191-199)
};
bool build_panel(const GfaGraph& g, Panel& p, std::string& err);

// ---------------------------------------------------------------------------------- anchors (anchors.cpp)
struct SketchResult {                // outputs of the device stages (dg_sketch_reads / dg_index_walks)
    std::vector<uint64_t> spectrum;          // distinct read hashes ascending; id = index (Sp_R, :534-546)
    std::vector<uint32_t> read_count;        // reads containing the hash (kmer_count, :711-732)
    std::vector<uint64_t> n_minimizers;      // per walk, before the join (:474)
    std::vector<uint64_t> hit_off;           // [n_walks+1]
    std::vector<uint32_t> hit_sid;           // spectrum id per hit, walk order
    std::vector<uint64_t> hit_vtx_off;       // [n_hits+1]
    std::vector<int32_t> hit_vtx;
};
struct Anchors {
    // Anchor_hits[id][h] = list of vertex lists, after the panel-sharing filter and the occurrence sort
    // (src/solver.cpp:590-663), stored flat: occ_off[(id*H + h) .. +1] indexes occurrences,
    // vtx_off[occ .. +1] indexes vtx.
    int32_t n_ids = 0, n_walks = 0;
    std::vector<int64_t> occ_off;
    std::vector<int64_t> vtx_off;
    std::vector<int32_t> vtx;
    std::vector<uint8_t> homo_bv;            // per id (src/solver.cpp:830-879)
    std::vector<int64_t> anchors_per_walk;   // "Number of Anchors" (:675-685)
    int64_t n_hom = 0, n_het = 0;
    std::string fit_line;                    // the "[M::...] Fitted model" text after the prefix
};
void build_anchors(const Panel& p, const SketchResult& s, float threshold, int threads, Anchors& a);

// ---------------------------------------------------------------------------------- classifier (kgfit.cpp)
struct KGParams {                    // src/Classifier.hpp:18-34
    double zp_copy = 1.3, zp_copy_het = 1.3, u_v = 4.0, sd_v = 1.2, var_w = 2.0, p_d = 0.5;
    int max_copy = 5;
    double amb_margin = 0.05, p_e = 0.01, err_shape = 2.0;
    bool treat_error_as_ambiguous = true;
    double reject_cost = 0.4;
};
struct KGFit { KGParams P; double nll = 0; int valley_x = 0, peak_x = 0; };
// KGFitterBO::fit with the options Solver sets (src/solver.cpp:777-785; src/Fitter.hpp:207-408, grid branch)
KGFit kg_fit(const std::vector<std::pair<int, double>>& hist /* (multiplicity, freq) ascending */, int max_multiplicity,
             int threads);
// KmerGenieDiploidLike::classify(x).label == HOM (src/Classifier.hpp:59-80)
bool kg_is_hom(const KGParams& P, int x);

// ---------------------------------------------------------------------------------- expanded graph (expand.cpp)
struct ExpGraph {                    // src/ExpandedGraph.hpp:16-26
    std::vector<std::vector<std::pair<int32_t, int32_t>>> adj;   // (v, weight)
    std::vector<std::vector<int32_t>> color;
    std::vector<std::vector<int32_t>> original_vertex;
    std::vector<int32_t> haplotype;
    std::vector<int32_t> level;
    std::vector<std::vector<int32_t>> vertices_in_level;
};
struct AnchorRec {                   // src/approximator.h:11-18
    int startOrg, endOrg, startExp, endExp;
    std::vector<int> colours;
    int nodeID;
};
struct Expanded {
    ExpGraph g;                      // after topologically_reorder (src/ExpandedGraph.hpp:29-102)
    std::vector<uint8_t> color_homo_bv;
    std::vector<std::vector<AnchorRec>> anchorsByHap;
    int n_colours = 0;
};
// Approximator::solve up to g.topologically_reorder(sink) (src/approximator.cpp:1014-1256) and
// color_homo_bv (:1283-1290)
bool expand_graph(const Panel& p, const Anchors& a, Expanded& e, std::string& err);
// ExpandedGraph::strict_bfs_levelize_and_reorder (src/ExpandedGraph.hpp:269-409)
bool levelize(ExpGraph& g, std::string& err);

// ---------------------------------------------------------------------------------- reconstruction (reconstruct.cpp)
struct DiploidSolution { int r1 = 0, r2 = 0; std::string hap1, hap2; };
// path recovery of diploid_dp_approximation_solver (src/approximator.cpp:779-925)
bool stitch_diploid(const Panel& p, const ExpGraph& g, const std::vector<std::pair<int, int>>& p1_edges,
                    const std::vector<std::pair<int, int>>& p2_edges, DiploidSolution& out, std::string& err);
// best_r rule + path -> original vertices of dp_approximation_solver (src/approximator.cpp:116-167)
int haploid_best_r(const std::vector<int>& colours_by_r, std::string& log);
std::string haploid_sequence(const Panel& p, const ExpGraph& g, const std::vector<int32_t>& path);
// FASTA writers (src/approximator.cpp:1271-1277, :1314-1325)
bool write_fasta_haploid(const std::string& path, const std::string& seq);
bool write_fasta_diploid(const std::string& path, const std::string& s1, const std::string& s2);

}  // namespace dgh
