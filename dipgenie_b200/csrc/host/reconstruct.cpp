// reconstruct.cpp — from DP results back to sequences (host glue, a10/a11 of SURVEY 8a):
//   stitch_diploid   : the path-recovery block of diploid_dp_approximation_solver
//                      (reference src/approximator.cpp:732-755, :779-925)
//   haploid_best_r   : the "best r" angle rule (src/approximator.cpp:115-138)
//   haploid_sequence : path -> original vertices -> first-occurrence de-dup -> sequence (:153-167, :30-40, :1264-1267)
//   FASTA writers    : src/approximator.cpp:1271-1277, :1314-1325
#include <cmath>
#include <fstream>
#include <queue>
#include <sstream>
#include <unordered_set>

#include "dgh.h"

namespace dgh {

namespace {

int find_next_zero_hap(const ExpGraph& g, int src, int target_hap) {      // :732-755
    if (g.haplotype.at(src) == target_hap && !g.original_vertex.at(src).empty()) return src;
    std::queue<int> q;
    std::unordered_set<int> visited;
    q.push(src); visited.insert(src);
    while (!q.empty()) {
        int u = q.front(); q.pop();
        for (const auto& e : g.adj[u]) {
            if (e.second != 0) continue;
            if (!visited.insert(e.first).second) continue;
            if (g.haplotype.at(e.first) == target_hap && !g.original_vertex.at(e.first).empty()) return e.first;
            q.push(e.first);
        }
    }
    return -1;
}

bool stitch_one(const Panel& p, const ExpGraph& g, const std::vector<std::pair<int, int>>& edges, const char* tag,
                std::string& out, std::string& err) {
    const int L = (int)g.vertices_in_level.size();
    const int source = g.vertices_in_level.at(0).at(0);
    int start_exp = source;
    for (int i = 0; i < (int)edges.size(); i++) {
        const auto& edge = edges[i];
        if (g.original_vertex[edge.first].size() != 1) {
            std::ostringstream os;
            os << tag << ": Vertex " << edge.first << " in map back has " << g.original_vertex[edge.first].size() << " original vertices";
            err = os.str();
            return false;                                      // reference: exit(1) (:796-800)
        }
        const int end_exp = edge.first;
        const int h = g.haplotype.at(end_exp);
        if (h < 0 || h >= (int)p.paths.size()) { err = std::string(tag) + ": recorded edge does not start on a haplotype lane"; return false; }
        if (start_exp == source && L > 1)
            for (int v : g.vertices_in_level.at(1)) if (g.haplotype.at(v) == h) start_exp = v;
        const int start_org = g.original_vertex.at(start_exp).at(0);
        const int end_org = g.original_vertex.at(end_exp).at(0);
        bool activated = false;
        for (int t = 0; t < (int)p.paths[h].size(); t++) {
            if (p.paths[h][t] == start_org) activated = true;
            if (activated) out += p.node_seq[p.paths[h][t]];
            if (p.paths[h][t] == end_org) { activated = false; break; }
        }
        if (g.level.at(edge.second) == L - 1) break;
        if (i + 1 >= (int)edges.size()) { err = std::string(tag) + ": edge list ends before the sink level"; return false; }
        const int next_hap = g.haplotype.at(edges[i + 1].first);
        const int next_start = find_next_zero_hap(g, edge.second, next_hap);
        if (next_start != -1) start_exp = next_start;
    }
    return true;
}

}  // namespace

bool stitch_diploid(const Panel& p, const ExpGraph& g, const std::vector<std::pair<int, int>>& p1_edges,
                    const std::vector<std::pair<int, int>>& p2_edges, DiploidSolution& out, std::string& err) {
    out = DiploidSolution();
    out.r1 = (int)p1_edges.size() - 1;                        // :784-785
    out.r2 = (int)p2_edges.size() - 1;
    if (!stitch_one(p, g, p1_edges, "P1", out.hap1, err)) return false;
    return stitch_one(p, g, p2_edges, "P2", out.hap2, err);
}

int haploid_best_r(const std::vector<int>& colors_by_r, std::string& log) {
    std::ostringstream os;
    int best_r = 0;
    double max_delta = 0;
    for (size_t i = 0; i + 1 < colors_by_r.size(); ++i) {
        os << "r: " << i << " true score: " << colors_by_r[i] << "\n";
        int delta = colors_by_r[i + 1] - colors_by_r[i];
        if (std::abs(delta) > max_delta) max_delta = std::abs(delta);
    }
    for (size_t r = 0; r + 1 < colors_by_r.size(); ++r) {
        int delta = colors_by_r[r + 1] - colors_by_r[r];
        double angle_rad = std::atan(static_cast<double>(delta) / max_delta);
        double angle_deg = angle_rad * 180.0 / M_PI;
        os << "r: " << r << " -> " << r + 1 << ", \xCE\x94" << "colors: " << delta << ", angle: " << angle_deg << "\xC2\xB0" << "\n";
        if (angle_deg < 5) { best_r = (int)r; break; }       // HAP_ANGLE_THRESHOLD (:24)
    }
    log = os.str();
    return best_r;
}

std::string haploid_sequence(const Panel& p, const ExpGraph& g, const std::vector<int32_t>& path) {
    std::unordered_set<int> seen;
    std::string out;
    for (int32_t u : path)
        for (int32_t o : g.original_vertex[u])
            if (seen.insert(o).second) out += p.node_seq[o];
    return out;
}

bool write_fasta_haploid(const std::string& path, const std::string& seq) {
    std::ofstream f(path, std::ios::out);
    if (!f) return false;
    f << ">" << "dp_sol" << " LN:" << seq.size() << std::endl;
    for (size_t i = 0; i < seq.size(); i += 80) f << seq.substr(i, 80) << std::endl;
    return true;
}

bool write_fasta_diploid(const std::string& path, const std::string& s1, const std::string& s2) {
    std::ofstream f(path, std::ios::out);
    if (!f) return false;
    f << ">" << "sol_1" << " bp:" << s1.size() << std::endl;
    for (size_t i = 0; i < s1.size(); i += 80) f << s1.substr(i, 80) << std::endl;
    f << ">" << "sol_2" << " bp:" << s2.size() << std::endl;
    for (size_t i = 0; i < s2.size(); i += 80) f << s2.substr(i, 80) << std::endl;
    return true;
}

}  // namespace dgh
