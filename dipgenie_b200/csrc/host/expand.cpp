// expand.cpp — the haplotype-expanded graph the DP runs on (host glue; defines the DP's input, so every
// creation order below is the reference's):
//   expand_graph : Approximator::solve up to g.topologically_reorder(sink) (reference
//                  src/approximator.cpp:1014-1256), plus color_homo_bv (:1283-1290)
//   kahn_reorder : ExpandedGraph::topologically_reorder (src/ExpandedGraph.hpp:29-102)
//   levelize     : ExpandedGraph::strict_bfs_levelize_and_reorder (src/ExpandedGraph.hpp:269-409)
#include <algorithm>
#include <numeric>
#include <queue>

#include "dgh.h"

namespace dgh {

namespace {

bool kahn_reorder(ExpGraph& g, int sink, std::string& err) {
    const size_t n = g.adj.size();
    std::vector<int> indeg(n, 0);
    for (auto& nb : g.adj) for (auto& e : nb) ++indeg[e.first];
    std::queue<int> q;
    for (size_t v = 0; v < n; ++v) if (indeg[v] == 0 && (int)v != sink) q.push((int)v);   // the sink is never queued
    bool sink_ready = indeg[sink] == 0;
    std::vector<int> order;
    order.reserve(n);
    while (!q.empty() || sink_ready) {
        int u;
        if (!q.empty()) { u = q.front(); q.pop(); }
        else { u = sink; sink_ready = false; }                                           // queue empty: only the sink is left
        order.push_back(u);
        for (auto& e : g.adj[u])
            if (--indeg[e.first] == 0) { if (e.first == sink) sink_ready = true; else q.push(e.first); }
    }
    if (order.size() != n) { err = "Graph contains a cycle; topological order impossible"; return false; }
    std::vector<int> new_idx(n);
    for (size_t i = 0; i < n; ++i) new_idx[order[i]] = (int)i;
    std::vector<std::vector<int32_t>> nc(n), no(n);
    std::vector<int32_t> nh(n);
    for (size_t i = 0; i < n; ++i) { nc[i] = std::move(g.color[order[i]]); no[i] = std::move(g.original_vertex[order[i]]); nh[i] = g.haplotype[order[i]]; }
    g.color.swap(nc); g.original_vertex.swap(no); g.haplotype.swap(nh);
    std::vector<std::vector<std::pair<int32_t, int32_t>>> na(n);
    for (size_t u = 0; u < n; ++u) {              // the lists move, their targets are renamed in place
        for (auto& e : g.adj[u]) e.first = new_idx[e.first];
        na[new_idx[u]] = std::move(g.adj[u]);
    }
    g.adj.swap(na);
    return true;
}

}  // namespace

bool expand_graph(const Panel& p, const Anchors& a, Expanded& ex, std::string& err) {
    ex = Expanded();
    const int H = (int)p.paths.size();
    const int32_t n_vtx = p.n_vtx;
    int64_t total = 0;
    for (auto& pw : p.paths) total += (int64_t)pw.size();
    auto& adj = ex.g.adj;
    auto& orig = ex.g.original_vertex;
    auto& hap = ex.g.haplotype;
    adj.assign((size_t)total + 2, {});
    orig.assign((size_t)total + 2, {});
    hap.assign((size_t)total + 2, 0);                 // source and sink keep haplotype 0 like the reference's value-initialised vector
    std::vector<int32_t> v2e((size_t)n_vtx * (size_t)std::max(H, 1), -1);   // vertex_to_expanded_map[v][h]
    auto V2E = [&](int32_t v, int h) -> int32_t& { return v2e[(size_t)v * H + h]; };
    const int sink = (int)adj.size() - 1;

    // lanes (:1028-1049)
    int32_t cur = 1;
    for (int h = 0; h < H; ++h) {
        adj[0].push_back({cur, 0});
        const auto& pw = p.paths[h];
        for (size_t i = 0; i < pw.size(); ++i) {
            V2E(pw[i], h) = cur;
            orig[cur].push_back(pw[i]);
            hap[cur] = h;
            adj[cur].push_back({i + 1 < pw.size() ? cur + 1 : sink, 0});
            ++cur;
        }
    }

    // recombination vertices (:1051-1095): one per original edge (u, adjacency slot j) that is off-path for some walk
    std::vector<std::vector<int32_t>> w_uv(n_vtx);
    for (int32_t u = 0; u < n_vtx; ++u) w_uv[u].assign(p.adj[u].size(), -1);
    cur = (int32_t)adj.size();
    for (int h = 0; h < H; ++h) {
        const auto& pw = p.paths[h];
        for (size_t i = 0; i < pw.size(); ++i) {
            const int32_t u = pw[i];
            for (size_t j = 0; j < p.adj[u].size(); ++j) {
                const int32_t v = p.adj[u][j];
                if (i + 1 == pw.size() || v != pw[i + 1]) {
                    if (w_uv[u][j] == -1) {
                        adj.emplace_back(); orig.emplace_back(); hap.push_back(-1);
                        w_uv[u][j] = cur++;
                    }
                    const int32_t rv = w_uv[u][j];
                    adj[V2E(u, h)].push_back({rv, 1});
                    if (adj[rv].empty())
                        for (int h2 = 0; h2 < H; ++h2) { const int32_t ve = V2E(v, h2); if (ve >= 0) adj[rv].push_back({ve, 0}); }
                }
            }
        }
    }

    // anchors -> colours and super-nodes (:1114-1176)
    auto& color = ex.g.color;
    color.assign(adj.size(), {});
    ex.anchorsByHap.assign(H, {});
    std::vector<int32_t> color_to_anchor;
    int nextID = (int)adj.size(), colourID = 0;
    for (int32_t id = 0; id < a.n_ids; ++id) {
        bool used = false;
        for (int h = 0; h < H; ++h) {
            for (int64_t o = a.occ_off[(size_t)id * H + h]; o < a.occ_off[(size_t)id * H + h + 1]; ++o) {
                const int64_t b = a.vtx_off[o], e = a.vtx_off[o + 1];
                if (b == e) continue;
                used = true;
                const int startOrig = a.vtx[b], endOrig = a.vtx[e - 1];
                const int startExp = V2E(startOrig, h), endExp = V2E(endOrig, h);
                int nodeID;
                if (startExp == endExp) nodeID = startExp;
                else {
                    adj[startExp].push_back({nextID, 0});
                    adj.push_back({{endExp, 0}});
                    orig.emplace_back(a.vtx.begin() + b, a.vtx.begin() + e);
                    color.emplace_back();
                    hap.push_back(-1);
                    nodeID = nextID++;
                }
                ex.anchorsByHap[h].push_back({startOrig, endOrig, startExp, endExp, {colourID}, nodeID});
            }
        }
        if (used) { color_to_anchor.push_back(id); ++colourID; }
    }
    ex.n_colours = colourID;

    // per-walk interval sweep (:1193-1246): overlap edges and colour propagation through containment.
    // std::sort on purpose (not stable_sort): the reference's order among equal (startExp,endExp) keys is
    // whatever libstdc++'s introsort produces for this comparator on this input order.
    for (int h = 0; h < H; ++h) {
        auto& vec = ex.anchorsByHap[h];
        if (vec.empty()) continue;
        std::sort(vec.begin(), vec.end(), [](const AnchorRec& x, const AnchorRec& y) {
            if (x.startExp != y.startExp) return x.startExp < y.startExp;
            return x.endExp < y.endExp;
        });
        std::vector<AnchorRec*> stk;
        for (auto& anc : vec) {
            while (!stk.empty() && stk.back()->endExp < anc.startExp) stk.pop_back();
            if (!stk.empty() && anc.startExp <= stk.back()->endExp && stk.back()->nodeID != anc.nodeID)
                adj[stk.back()->nodeID].push_back({anc.nodeID, 0});
            for (int i = (int)stk.size() - 1; i >= 0; --i) {
                if (anc.endExp <= stk[i]->endExp) {
                    for (int c : anc.colours)
                        if (std::find(stk[i]->colours.begin(), stk[i]->colours.end(), c) == stk[i]->colours.end())
                            stk[i]->colours.push_back(c);
                } else break;
            }
            stk.push_back(&anc);
        }
        for (const auto& anc : vec) {
            auto& dst = color[anc.nodeID];
            dst.insert(dst.end(), anc.colours.begin(), anc.colours.end());
            std::sort(dst.begin(), dst.end());
            dst.erase(std::unique(dst.begin(), dst.end()), dst.end());
        }
    }

    // color_homo_bv (:1283-1290)
    ex.color_homo_bv.assign((size_t)colourID, 0);
    for (int c = 0; c < colourID; ++c) ex.color_homo_bv[c] = a.homo_bv[color_to_anchor[c]] ? 1 : 0;

    return kahn_reorder(ex.g, sink, err);
}

bool levelize(ExpGraph& g, std::string& err) {
    const int n0 = (int)g.adj.size();
    if (n0 == 0) return true;
    std::vector<int> indeg(n0, 0), outdeg(n0, 0);
    for (int u = 0; u < n0; ++u) { outdeg[u] = (int)g.adj[u].size(); for (auto& e : g.adj[u]) ++indeg[e.first]; }
    int source = -1;
    for (int v = 0; v < n0; ++v)
        if (indeg[v] == 0 && outdeg[v] > 0) {
            if (source == -1) source = v;
            else { err = "Uh oh, multiple potential sources found while leveling"; return false; }   // reference: exit(-1)
        }
    if (source < 0) { err = "bad source index"; return false; }
    // 1) BFS distance from the source (:292-301)
    std::vector<int> dist(n0, -1);
    {
        std::queue<int> q;
        dist[source] = 0; q.push(source);
        while (!q.empty()) {
            int u = q.front(); q.pop();
            for (auto& e : g.adj[u]) if (dist[e.first] == -1) { dist[e.first] = dist[u] + 1; q.push(e.first); }
        }
    }
    // 2) topological order (:303-312)
    std::vector<int> topo;
    topo.reserve(n0);
    {
        std::queue<int> q;
        for (int v = 0; v < n0; ++v) if (indeg[v] == 0) q.push(v);
        while (!q.empty()) {
            int u = q.front(); q.pop();
            topo.push_back(u);
            for (auto& e : g.adj[u]) if (--indeg[e.first] == 0) q.push(e.first);
        }
        if ((int)topo.size() != n0) { err = "Graph contains a cycle; strict leveling requires a DAG"; return false; }
    }
    // 3) longest-path relaxation seeded with the BFS distance (:314-317)
    std::vector<int> lvl(n0, 0);
    for (int v = 0; v < n0; ++v) if (dist[v] >= 0) lvl[v] = dist[v];
    for (int u : topo) for (auto& e : g.adj[u]) if (lvl[e.first] <= lvl[u]) lvl[e.first] = lvl[u] + 1;
    // 4) dummy vertices so that every edge spans exactly one level (:319-352)
    std::vector<std::vector<std::pair<int32_t, int32_t>>> nadj(n0);
    std::vector<int32_t> nlvl(lvl.begin(), lvl.end());
    {
        size_t n_dummy = 0;                         // (known up front: no regrowth of the per-vertex tables)
        for (int u = 0; u < n0; ++u)
            for (auto& e : g.adj[u]) { const int gap = lvl[e.first] - lvl[u] - 1; if (gap > 0) n_dummy += (size_t)gap; }
        nadj.reserve((size_t)n0 + n_dummy); nlvl.reserve((size_t)n0 + n_dummy);
        g.color.reserve((size_t)n0 + n_dummy); g.original_vertex.reserve((size_t)n0 + n_dummy); g.haplotype.reserve((size_t)n0 + n_dummy);
    }
    for (int u = 0; u < n0; ++u) {
        nadj[u].reserve(g.adj[u].size());
        for (auto& e : g.adj[u]) {
            const int v = e.first, w = e.second;
            const int gap = nlvl[v] - nlvl[u] - 1;
            if (gap <= 0) { nadj[u].emplace_back(v, w); continue; }
            int prev = u;
            for (int step = 1; step <= gap; ++step) {
                const int dummy = (int)nadj.size();
                nadj.emplace_back();
                g.color.emplace_back();
                { std::vector<int32_t> inherited = g.original_vertex[u]; g.original_vertex.push_back(std::move(inherited)); }
                nlvl.push_back(nlvl[u] + step);
                g.haplotype.push_back(g.haplotype[u]);
                nadj[prev].emplace_back(dummy, step == 1 ? w : 0);
                prev = dummy;
            }
            nadj[prev].emplace_back(v, 0);
        }
    }
    g.adj.swap(nadj);
    // 5) order by (level, id) (:360-400)
    const int n1 = (int)g.adj.size();
    int max_level = 0;
    for (int v = 0; v < n1; ++v) max_level = std::max(max_level, nlvl[v]);
    std::vector<int> order(n1);                     // (level, id) ascending: a counting sort by level, ids stay in order
    std::vector<int> lvl_start((size_t)max_level + 2, 0);
    for (int v = 0; v < n1; ++v) ++lvl_start[(size_t)nlvl[v] + 1];
    for (int l = 0; l <= max_level; ++l) lvl_start[(size_t)l + 1] += lvl_start[l];
    {
        std::vector<int> cursor(lvl_start.begin(), lvl_start.end() - 1);
        for (int v = 0; v < n1; ++v) order[(size_t)cursor[nlvl[v]]++] = v;
    }
    std::vector<int> new_id(n1);
    for (int i = 0; i < n1; ++i) new_id[order[i]] = i;
    std::vector<std::vector<int32_t>> nc(n1), no(n1);
    std::vector<int32_t> nl(n1), nh(n1);
    for (int i = 0; i < n1; ++i) {
        const int old = order[i];
        nc[i] = std::move(g.color[old]); no[i] = std::move(g.original_vertex[old]); nl[i] = nlvl[old]; nh[i] = g.haplotype[old];
    }
    g.color.swap(nc); g.original_vertex.swap(no); g.level.swap(nl); g.haplotype.swap(nh);
    std::vector<std::vector<std::pair<int32_t, int32_t>>> na(n1);
    for (int u = 0; u < n1; ++u) {
        for (auto& e : g.adj[u]) e.first = new_id[e.first];
        na[new_id[u]] = std::move(g.adj[u]);
    }
    g.adj.swap(na);
    // 7) per-level buckets (:402-407)
    g.vertices_in_level.assign(max_level + 1, {});
    for (int l = 0; l <= max_level; ++l) {          // vertex ids are (level, id)-ordered now: level l owns a contiguous range
        auto& b = g.vertices_in_level[l];
        b.resize((size_t)(lvl_start[(size_t)l + 1] - lvl_start[l]));
        std::iota(b.begin(), b.end(), lvl_start[l]);
    }
    return true;
}

}  // namespace dgh
