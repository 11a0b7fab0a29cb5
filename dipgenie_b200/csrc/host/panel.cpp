// panel.cpp — the panel model the hot path works on: Solver::read_gfa (reference src/solver.cpp:27-227)
// and Solver::read_ip_reads (:230-245), written for this repo.
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <limits>

#include "dgh.h"

namespace dgh {

bool build_panel(const GfaGraph& g, Panel& p, std::string& err) {
    p = Panel();
    const int32_t V = (int32_t)g.seg_name.size();
    p.n_vtx = V;
    p.node_seq = g.seg_seq;                                           // :37-43
    // forward-strand adjacency with the target's orientation dropped (:60-91)
    p.adj.assign(V, {});
    for (const auto& a : g.arcs)
        if ((a.first & 1) == 0) p.adj[a.first >> 1].push_back((int32_t)(a.second >> 1));
    // walks (:108-125): any reverse-strand step aborts, like the reference's silent exit(1) (:116-119)
    const size_t H = g.walks.size();
    p.paths.assign(H, {});
    for (size_t w = 0; w < H; ++w) {
        p.walk_names.push_back(g.walks[w].sample + "." + std::to_string(g.walks[w].hap));
        p.paths[w].reserve(g.walks[w].v.size());
        for (uint32_t x : g.walks[w].v) {
            if (x & 1) { err = "walk " + p.walk_names.back() + " visits a reverse-strand vertex (reference exits here, solver.cpp:116-119)"; return false; }
            p.paths[w].push_back((int32_t)(x >> 1));
        }
    }
    // MSA-like column of every vertex (:127-171): earliest walk position, then monotone relaxation along walks
    const int64_t INF = std::numeric_limits<int64_t>::max() / 4;
    std::vector<int64_t> pos(V, INF);
    for (const auto& pw : p.paths)
        for (int64_t t = 0; t < (int64_t)pw.size(); ++t)
            if (t < pos[pw[t]]) pos[pw[t]] = t;
    int64_t max_seed = -1;
    for (int32_t v = 0; v < V; ++v) if (pos[v] != INF) max_seed = std::max(max_seed, pos[v]);
    const int64_t fallback = max_seed >= 0 ? max_seed + 1 : 0;
    for (int32_t v = 0; v < V; ++v) if (pos[v] == INF) pos[v] = fallback;
    // The reference sweeps the walks until nothing moves (:150-171): the least solution of pos[next] >= pos[prev] + 1 above
    // the seeds.  On an acyclic panel that is one longest-path pass over the walk steps in topological order (the sweeps took
    // 0.8 s of a 2-s run on MHC_4: one per link of the longest chain of corrections); a cycle among the walk steps falls
    // back to the reference's capped sweeps, whose result then depends on the sweep order.
    {
        std::vector<int32_t> indeg(V, 0), head((size_t)V + 1, 0);
        size_t n_steps = 0;
        for (const auto& pw : p.paths) n_steps += pw.empty() ? 0 : pw.size() - 1;
        for (const auto& pw : p.paths)
            for (size_t t = 1; t < pw.size(); ++t) { ++head[(size_t)pw[t - 1] + 1]; ++indeg[pw[t]]; }
        for (int32_t v = 0; v < V; ++v) head[(size_t)v + 1] += head[v];
        std::vector<int32_t> succ(n_steps), fill(head.begin(), head.end() - 1), order;
        for (const auto& pw : p.paths)
            for (size_t t = 1; t < pw.size(); ++t) succ[(size_t)fill[pw[t - 1]]++] = pw[t];
        order.reserve(V);
        for (int32_t v = 0; v < V; ++v) if (indeg[v] == 0) order.push_back(v);
        std::vector<int64_t> lp(pos);
        for (size_t x = 0; x < order.size(); ++x) {
            const int32_t u = order[x];
            for (int32_t e = head[u]; e < head[(size_t)u + 1]; ++e) {
                const int32_t v = succ[(size_t)e];
                if (lp[v] < lp[u] + 1) lp[v] = lp[u] + 1;
                if (--indeg[v] == 0) order.push_back(v);
            }
        }
        if ((int32_t)order.size() == V) pos.swap(lp);
        else {
            bool changed = true;
            int iter = 0;
            const int iter_cap = std::max(10, V);
            while (changed && iter++ < iter_cap) {
                changed = false;
                for (const auto& pw : p.paths)
                    for (size_t t = 1; t < pw.size(); ++t) {
                        const int64_t need = pos[pw[t - 1]] + 1;
                        if (pos[pw[t]] < need) { pos[pw[t]] = need; changed = true; }
                    }
            }
        }
    }
    // dense ranks (:173-199)
    std::vector<std::pair<int64_t, int32_t>> by_pos;
    by_pos.reserve(V);
    for (int32_t v = 0; v < V; ++v) by_pos.emplace_back(pos[v], v);
    std::sort(by_pos.begin(), by_pos.end());
    std::vector<int32_t> dense(V, -1);
    int32_t rank = -1;
    int64_t prev = std::numeric_limits<int64_t>::min();
    p.top_order_map.assign(V, -1);
    for (int32_t i = 0; i < V; ++i) {
        if (by_pos[i].first != prev) { ++rank; prev = by_pos[i].first; }
        dense[by_pos[i].second] = rank;
        p.top_order_map[by_pos[i].second] = i;
    }
    // adjacency order = (column, id) (:216-223); this fixes the creation order of recombination vertices
    for (int32_t u = 0; u < V; ++u)
        std::sort(p.adj[u].begin(), p.adj[u].end(), [&](int32_t a, int32_t b) {
            if (dense[a] != dense[b]) return dense[a] < dense[b];
            return a < b;
        });
    return true;
}

namespace {
struct ByteReader {
    gzFile fp;
    std::vector<unsigned char> buf;
    size_t pos = 0, len = 0;
    bool eof = false;
    explicit ByteReader(gzFile f) : fp(f), buf(1 << 20) {}
    int getc() {
        if (pos == len) {
            if (eof) return -1;
            int n = gzread(fp, buf.data(), (unsigned)buf.size());
            if (n <= 0) { eof = true; return -1; }
            pos = 0; len = (size_t)n;
        }
        return buf[pos++];
    }
    // appends the rest of the current line (without the line break) to s; returns false at EOF with nothing read
    bool rest_of_line(std::string* s) {
        bool any = false;
        for (;;) {
            int c = getc();
            if (c < 0) return any;
            any = true;
            if (c == '\n') break;
            if (s) s->push_back((char)c);
        }
        if (s && !s->empty() && s->back() == '\r') s->pop_back();
        return true;
    }
};
}  // namespace

// kseq_read semantics (src/kseq.h:183-233): '>' or '@' starts a record; the name ends at the first white
// space; sequence lines run until a line starting with '>', '+' or '@'; for FASTQ the quality is read until
// it is at least as long as the sequence; a length mismatch ends the file (kseq returns -2, the reference's
// loop `while ((l = kseq_read(seq)) >= 0)` stops).
bool read_sequences(const std::string& path, std::vector<std::string>& seqs, std::string& err) {
    gzFile fp = gzopen(path.c_str(), "r");
    if (!fp) { err = "cannot open " + path; return false; }
    gzbuffer(fp, 1 << 20);
    ByteReader br(fp);
    int last = 0;
    for (;;) {
        int c;
        if (last == 0) {
            while ((c = br.getc()) >= 0 && c != '>' && c != '@') {}
            if (c < 0) break;
        }
        last = 0;
        br.rest_of_line(nullptr);                       // name + comment (unused downstream)
        std::string seq;
        while ((c = br.getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq.push_back((char)c);
            br.rest_of_line(&seq);
        }
        if (c == '>' || c == '@') last = c;
        if (c != '+') { seqs.push_back(std::move(seq)); if (c < 0) break; continue; }
        br.rest_of_line(nullptr);                       // rest of the '+' line
        std::string qual;
        while (qual.size() < seq.size()) if (!br.rest_of_line(&qual)) break;
        if (qual.size() != seq.size()) break;           // kseq error -2: reading stops, record dropped
        seqs.push_back(std::move(seq));
    }
    gzclose(fp);
    return true;
}

}  // namespace dgh
