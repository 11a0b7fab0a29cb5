// gfa.cpp — GFA 1.1 reader of the host front end (own code; behaviour follows the reference's
// gfatools-derived parser so that segment ids, arcs and walks come out identical):
//   * segment id = order of first mention on an S or L line (src/gfa-base.cpp:75-96, gfa_add_seg);
//   * S: name, sequence ('*' = absent, then LN:i: gives the length) (src/gfa-io.cpp:214-283);
//   * L: v, orientation, w, orientation, overlap ('*', CIGAR or a:b) (src/gfa-io.cpp:285-367);
//   * W: sample, hap index, contig, start, end, walk string of >seg / <seg steps; unknown segments are
//     skipped (src/gfa-io.cpp:369-432);
//   * after the whole file: walks flipped to their majority strand (gfa_walk_flip, src/gfa-io.cpp:64-114),
//     then gfa_finalize (src/gfa-base.cpp:421-430): segments without length are deleted, every arc gets
//     its complement arc unless an equal one exists, arcs touching deleted segments are dropped.
//   Lines that are not S/L/W, tags other than LN, and the FASTA-in-GFA mode are ignored.
#include <zlib.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "dgh.h"

namespace dgh {

namespace {

struct Arc {
    uint32_t v, w;
    int32_t ov, ow;
    bool comp = false, del = false;
};

struct LineReader {
    gzFile fp = nullptr;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool eof = false;
    explicit LineReader(gzFile f) : fp(f), buf(1 << 20) {}
    bool next(std::string& line) {
        line.clear();
        for (;;) {
            if (pos == len) {
                if (eof) return !line.empty();
                int n = gzread(fp, buf.data(), (unsigned)buf.size());
                if (n <= 0) { eof = true; return !line.empty(); }
                pos = 0; len = (size_t)n;
            }
            const char* s = buf.data() + pos;
            const char* nl = (const char*)memchr(s, '\n', len - pos);
            if (nl) {
                line.append(s, nl - s);
                pos += (size_t)(nl - s) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(s, len - pos);
            pos = len;
        }
    }
};

inline void split_tabs(const std::string& s, std::vector<std::pair<size_t, size_t>>& f, size_t max_fields) {
    f.clear();
    size_t b = 0;
    while (f.size() + 1 < max_fields) {
        size_t e = s.find('\t', b);
        if (e == std::string::npos) break;
        f.emplace_back(b, e - b);
        b = e + 1;
    }
    f.emplace_back(b, s.size() - b);
}

// overlap field of an L line (src/gfa-io.cpp:311-337)
void parse_overlap(const char* q, int32_t& ov, int32_t& ow) {
    ov = ow = 0;
    if (*q == '*' || *q == 0) return;
    if (*q == ':') { ov = INT32_MAX; ow = isdigit((unsigned char)q[1]) ? (int32_t)strtol(q + 1, nullptr, 10) : INT32_MAX; return; }
    if (!isdigit((unsigned char)*q)) return;
    char* r;
    long first = strtol(q, &r, 10);
    if (isupper((unsigned char)*r)) {          // CIGAR
        const char* c = q;
        ov = ow = 0;
        do {
            char* e;
            long l = strtol(c, &e, 10);
            if (*e == 'M' || *e == 'D' || *e == 'N') ov += (int32_t)l;
            if (*e == 'M' || *e == 'I' || *e == 'S') ow += (int32_t)l;
            c = e + 1;
        } while (isdigit((unsigned char)*c));
    } else if (*r == ':') {
        ov = (int32_t)first;
        ow = isdigit((unsigned char)r[1]) ? (int32_t)strtol(r + 1, nullptr, 10) : INT32_MAX;
    }
}

}  // namespace

bool read_gfa_file(const std::string& path, GfaGraph& g, std::string& err) {
    gzFile fp = gzopen(path.c_str(), "r");
    if (!fp) { err = "cannot open " + path; return false; }
    gzbuffer(fp, 1 << 20);
    g = GfaGraph();
    std::unordered_map<std::string, int32_t> ids;
    auto add_seg = [&](const std::string& name) -> int32_t {
        auto it = ids.find(name);
        if (it != ids.end()) return it->second;
        int32_t id = (int32_t)g.seg_name.size();
        ids.emplace(name, id);
        g.seg_name.push_back(name);
        g.seg_seq.emplace_back();
        g.seg_len.push_back(0);
        return id;
    };
    std::vector<Arc> arcs;
    LineReader lr(fp);
    std::string line;
    std::vector<std::pair<size_t, size_t>> f;
    while (lr.next(line)) {
        if (line.size() < 3 || line[1] != '\t') continue;
        const char t = line[0];
        if (t == 'S') {
            split_tabs(line, f, 64);
            if (f.size() < 3) continue;
            const std::string name = line.substr(f[1].first, f[1].second);
            int32_t LN = -1;
            for (size_t x = 3; x < f.size(); ++x)
                if (f[x].second > 5 && line.compare(f[x].first, 5, "LN:i:") == 0) LN = atoi(line.c_str() + f[x].first + 5);
            const bool has_seq = !(f[2].second >= 1 && line[f[2].first] == '*');
            const int32_t sid = add_seg(name);
            if (has_seq) { g.seg_seq[sid] = line.substr(f[2].first, f[2].second); g.seg_len[sid] = (int32_t)f[2].second; }
            else { g.seg_seq[sid].clear(); g.seg_len[sid] = LN >= 0 ? LN : 0; }
        } else if (t == 'L') {
            split_tabs(line, f, 7);
            if (f.size() < 5) continue;
            const char o1 = line[f[2].first], o2 = line[f[4].first];
            if ((o1 != '+' && o1 != '-') || (o2 != '+' && o2 != '-')) continue;
            Arc a;
            a.v = (uint32_t)add_seg(line.substr(f[1].first, f[1].second)) << 1 | (o1 != '+');
            a.w = (uint32_t)add_seg(line.substr(f[3].first, f[3].second)) << 1 | (o2 != '+');
            if (f.size() >= 6) { std::string ovs = line.substr(f[5].first, f[5].second); parse_overlap(ovs.c_str(), a.ov, a.ow); }
            else a.ov = a.ow = 0;
            arcs.push_back(a);
        } else if (t == 'W') {
            split_tabs(line, f, 8);
            if (f.size() < 7) continue;
            GfaWalk w;
            w.sample = line.substr(f[1].first, f[1].second);
            w.hap = atoi(line.c_str() + f[2].first);
            const size_t b = f[6].first, e = f[6].first + f[6].second;
            size_t q = b;
            while (q < e) {
                const char d = line[q];
                if (d != '>' && d != '<') { ++q; continue; }
                size_t r = q + 1;
                while (r < e && line[r] != '>' && line[r] != '<') ++r;
                auto it = ids.find(line.substr(q + 1, r - q - 1));
                // (the reference adds a sequence-less segment for an unknown name, src/gfa-io.cpp: gfa_add_seg, which its
                // finalize step then deletes; a walk through a segment without sequence is a malformed input here)
                if (it == ids.end()) {
                    err = "W line of " + w.sample + " visits segment " + line.substr(q + 1, r - q - 1) + ", which has no S line before it";
                    gzclose(fp);
                    return false;
                }
                w.v.push_back((uint32_t)it->second << 1 | (d == '<'));
                q = r;
            }
            g.walks.push_back(std::move(w));
        }
    }
    gzclose(fp);

    const int32_t n_seg = (int32_t)g.seg_name.size();
    // ---- gfa_walk_flip (src/gfa-io.cpp:64-114) ----
    if (!g.walks.empty()) {
        std::vector<int8_t> strand(n_seg, 0);
        for (auto& w : g.walks)
            for (uint32_t x : w.v)
                if (strand[x >> 1] == 0) strand[x >> 1] = (x & 1) ? -1 : 1;
        for (auto& w : g.walks) {
            int64_t agree = 0, dis = 0;
            for (uint32_t x : w.v) ((((x & 1) ? -1 : 1) == strand[x >> 1]) ? agree : dis)++;
            if (agree >= dis) continue;
            std::reverse(w.v.begin(), w.v.end());
            for (auto& x : w.v) x ^= 1;
        }
    }
    // ---- gfa_finalize (src/gfa-base.cpp:421-430) ----
    g.seg_del.assign(n_seg, 0);
    for (int32_t s = 0; s < n_seg; ++s) if (g.seg_len[s] == 0) g.seg_del[s] = 1;     // gfa_fix_no_seg
    std::stable_sort(arcs.begin(), arcs.end(), [](const Arc& a, const Arc& b) { return a.v < b.v; });
    const size_t n_orig = arcs.size();
    std::vector<size_t> first(2 * (size_t)n_seg + 1, 0);
    for (size_t i = 0; i < n_orig; ++i) ++first[arcs[i].v + 1];
    for (size_t v = 0; v < 2 * (size_t)n_seg; ++v) first[v + 1] += first[v];
    // gfa_fix_symm_add: pair every arc with its complement among the ORIGINAL arcs, else append one
    for (size_t i = 0; i < n_orig; ++i) {
        if (arcs[i].del || arcs[i].comp) continue;
        const uint32_t v = arcs[i].v, wc = arcs[i].w ^ 1;
        bool found = false;
        for (size_t j = first[wc]; j < first[wc + 1]; ++j) {
            Arc& b = arcs[j];
            if (b.del || b.comp) continue;
            if (b.w == (v ^ 1) && b.ov == arcs[i].ow && b.ow == arcs[i].ov) { b.comp = true; found = true; break; }
        }
        if (!found) {
            Arc c;
            c.v = wc; c.w = v ^ 1; c.ov = arcs[i].ow; c.ow = arcs[i].ov; c.comp = true;
            arcs.push_back(c);
        }
    }
    // gfa_fix_arc_len + gfa_cleanup: drop arcs touching deleted segments
    for (auto& a : arcs) if (g.seg_del[a.v >> 1] || g.seg_del[a.w >> 1]) a.del = true;
    for (auto& a : arcs) if (!a.del) g.arcs.emplace_back(a.v, a.w);
    return true;
}

}  // namespace dgh
