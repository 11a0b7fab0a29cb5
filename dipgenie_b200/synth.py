"""Seeded synthetic inputs for the DP kernels.

* `random_level_graph`  — arbitrary random levelized DAGs (ragged widths, multi-edges, high fan-in,
  weights 0/1, random colour sets).  Used by the parity tests: any levelized DAG is a legal input of
  the reference's diploid DP (src/approximator.cpp:362).
* `lane_panel_graph`    — a structural model of the haplotype-expanded graph the reference builds
  (src/approximator.cpp:1018-1246): H haplotype lanes joined by weight-0 lane edges, recombination
  vertices with weight-1 fan-in from a subset of lanes and weight-0 fan-out to every lane, anchors
  (colours) on lane vertices.  Used to size the sweep for wide panels (BASELINE config 4 shape)
  without running the whole front end.
"""
from __future__ import annotations

import numpy as np

from .cuda_api import LevelGraph


def _csr(lists_len, flat):
    off = np.zeros(len(lists_len) + 1, np.int64)
    np.cumsum(lists_len, out=off[1:])
    return off, np.asarray(flat)


def random_level_graph(seed: int, n_levels: int = 12, max_width: int = 6, n_colours: int = 10,
                       p_weight1: float = 0.3, p_colour: float = 0.4, max_out: int = 3,
                       p_dup: float = 0.1, hom_frac: float = 0.3) -> LevelGraph:
    rng = np.random.default_rng(seed)
    widths = [1] + [int(rng.integers(1, max_width + 1)) for _ in range(max(n_levels - 2, 0))] + ([1] if n_levels > 1 else [])
    level_off = np.concatenate([[0], np.cumsum(widths)]).astype(np.int32)
    V = int(level_off[-1])
    adj = [[] for _ in range(V)]
    for l in range(len(widths) - 1):
        lo, mid, hi = level_off[l], level_off[l + 1], level_off[l + 2]
        covered = set()
        for u in range(lo, mid):
            nout = int(rng.integers(1, max_out + 1))
            for _ in range(nout):
                v = int(rng.integers(mid, hi))
                w = int(rng.random() < p_weight1)
                adj[u].append((v, w))
                covered.add(v)
                if rng.random() < p_dup:
                    adj[u].append((v, w))          # duplicate same-weight edge
        for v in range(mid, hi):                    # every vertex reachable
            if v not in covered:
                u = int(rng.integers(lo, mid))
                adj[u].append((v, int(rng.random() < p_weight1)))
    # a (u,v) pair must not carry two different weights (holds for the reference's graphs)
    for u in range(V):
        seen = {}
        fixed = []
        for v, w in adj[u]:
            w = seen.setdefault(v, w)
            fixed.append((v, w))
        adj[u] = fixed
    outdeg = [len(a) for a in adj]
    adj_off, _ = _csr(outdeg, [])
    adj_dst = np.array([v for a in adj for v, _ in a], np.int32)
    adj_w = np.array([w for a in adj for _, w in a], np.uint8)
    cols = []
    for v in range(V):
        if n_colours > 0 and rng.random() < p_colour:
            n = int(rng.integers(1, min(n_colours, 5) + 1))
            cols.append(sorted(set(int(c) for c in rng.integers(0, n_colours, n))))
        else:
            cols.append([])
    col_off, _ = _csr([len(c) for c in cols], [])
    col_val = np.array([c for cs in cols for c in cs], np.int32)
    hom = (rng.random(max(n_colours, 0)) < hom_frac).astype(np.uint8)
    return LevelGraph(level_off, adj_off, adj_dst, adj_w, col_off, col_val, hom)


def lane_panel_graph(seed: int, n_lanes: int, n_blocks: int, rec_per_block: int = 2, fanin_frac: float = 0.3,
                     p_colour: float = 0.08, colours_per_vertex: int = 3, n_colours: int = 4096,
                     hom_frac: float = 0.15, run_len: int = 3) -> LevelGraph:
    """Levels come in blocks: `run_len` plain lane levels (width H), then a level that also holds
    `rec_per_block` recombination vertices (width H + rec) fed by weight-1 edges from a random
    `fanin_frac` of the lanes of the previous level and fanning out (weight 0) to every lane of the
    next level.  Source and sink are single vertices."""
    rng = np.random.default_rng(seed)
    H = n_lanes
    widths = [1]
    kinds = []                                   # per level after source: 'lane' or 'rec'
    for _ in range(n_blocks):
        for _ in range(run_len):
            widths.append(H); kinds.append("lane")
        widths.append(H + rec_per_block); kinds.append("rec")
    widths.append(H); kinds.append("lane")
    widths.append(1)
    level_off = np.concatenate([[0], np.cumsum(widths)]).astype(np.int32)
    V = int(level_off[-1])
    nl = len(widths)
    src_l, dst_l, w_l = [], [], []
    lanes = np.arange(H)
    # source -> lanes of level 1
    src_l.append(np.zeros(H, np.int64)); dst_l.append(level_off[1] + lanes); w_l.append(np.zeros(H, np.uint8))
    for l in range(1, nl - 2):
        lo, mid = level_off[l], level_off[l + 1]
        kind_here, kind_next = kinds[l - 1], kinds[l]
        # lane edges (lane vertices are the first H of every level)
        src_l.append(lo + lanes); dst_l.append(mid + lanes); w_l.append(np.zeros(H, np.uint8))
        if kind_next == "rec":                   # weight-1 fan-in into recombination vertices of level l+1
            for q in range(rec_per_block):
                m = max(1, int(round(fanin_frac * H)))
                sel = np.sort(rng.choice(H, m, replace=False))
                src_l.append(lo + sel); dst_l.append(np.full(m, mid + H + q)); w_l.append(np.ones(m, np.uint8))
        if kind_here == "rec":                   # weight-0 fan-out of this level's recombination vertices
            for q in range(rec_per_block):
                src_l.append(np.full(H, lo + H + q)); dst_l.append(mid + lanes); w_l.append(np.zeros(H, np.uint8))
    lo = level_off[nl - 2]
    src_l.append(lo + lanes); dst_l.append(np.full(H, level_off[nl - 1])); w_l.append(np.zeros(H, np.uint8))
    src = np.concatenate(src_l).astype(np.int64)
    dst = np.concatenate(dst_l).astype(np.int64)
    w = np.concatenate(w_l)
    order = np.lexsort((np.arange(len(src)), src))   # stable by source vertex
    src, dst, w = src[order], dst[order], w[order]
    outdeg = np.bincount(src, minlength=V)
    adj_off = np.concatenate([[0], np.cumsum(outdeg)]).astype(np.int64)
    # colours on lane vertices only
    is_lane = np.zeros(V, bool)
    for l in range(1, nl - 1):
        is_lane[level_off[l]: level_off[l] + H] = True
    coloured = is_lane & (rng.random(V) < p_colour)
    ncol = np.where(coloured, colours_per_vertex, 0)
    col_off = np.concatenate([[0], np.cumsum(ncol)]).astype(np.int64)
    col_val = rng.integers(0, n_colours, int(col_off[-1])).astype(np.int32)
    # keep each vertex's list sorted (the reference's colour lists are sorted sets)
    if len(col_val):
        vid = np.repeat(np.arange(V), ncol)
        o = np.lexsort((col_val, vid))
        col_val = col_val[o]
    hom = (rng.random(n_colours) < hom_frac).astype(np.uint8)
    return LevelGraph(level_off, adj_off, dst.astype(np.int32), w.astype(np.uint8), col_off, col_val, hom)


def truncate_levels(g: LevelGraph, n_levels: int) -> LevelGraph:
    """First `n_levels` levels of g with a single sink appended (every vertex of the last kept level gets one
    weight-0 edge to it) — a bounded sample of the same workload, still a legal levelized graph."""
    L = g.n_levels
    if n_levels >= L:
        return g
    V = int(g.level_off[n_levels])
    lo = int(g.level_off[n_levels - 1])
    outdeg = np.diff(g.adj_off)[:V].copy()
    e_keep = int(g.adj_off[lo])
    adj_dst = np.concatenate([g.adj_dst[:e_keep], np.full(V - lo, V, np.int32)])
    adj_w = np.concatenate([g.adj_w[:e_keep], np.zeros(V - lo, np.uint8)])
    outdeg[lo:V] = 1
    adj_off = np.concatenate([[0], np.cumsum(np.concatenate([outdeg, [0]]))]).astype(np.int64)
    col_off = np.concatenate([g.col_off[: V + 1], [g.col_off[V]]]).astype(np.int64)
    col_val = g.col_val[: int(g.col_off[V])]
    level_off = np.concatenate([g.level_off[: n_levels + 1], [V + 1]]).astype(np.int32)
    return LevelGraph(level_off, adj_off, adj_dst, adj_w, col_off, col_val, g.colour_is_hom)


def random_kahn_graph(seed: int, n: int = 200, max_out: int = 4, p_weight1: float = 0.3, p_colour: float = 0.3,
                      n_colours: int = 64, max_span: int = 12, max_cols: int = 4, p_dup: float = 0.1):
    """Random DAG whose vertex ids are already a topological order (what ExpandedGraph::topologically_reorder
    hands to the haploid DP): every vertex i < n-1 gets 1..max_out forward edges of span <= max_span (duplicates
    and mixed-weight parallel edges included, which exercise the first-writer tie-break), sorted colour sets."""
    from .cuda_api import HapGraph
    rng = np.random.default_rng(seed)
    off = [0]
    dst, wt = [], []
    for u in range(n):
        if u < n - 1:
            for _ in range(int(rng.integers(1, max_out + 1))):
                v = int(min(n - 1, u + 1 + rng.integers(0, max_span)))
                dst.append(v)
                wt.append(1 if rng.random() < p_weight1 else 0)
                if rng.random() < p_dup:
                    dst.append(v)
                    wt.append(1 if rng.random() < 0.5 else 0)
        off.append(len(dst))
    coff = [0]
    cval = []
    for u in range(n):
        if n_colours and rng.random() < p_colour:
            cs = np.unique(rng.integers(0, n_colours, int(rng.integers(1, max_cols + 1))))
            cval.extend(int(c) for c in cs)
        coff.append(len(cval))
    return HapGraph(np.array(off, np.int64), np.array(dst, np.int32), np.array(wt, np.uint8), np.array(coff, np.int64),
                    np.array(cval, np.int32), n_colours=max(n_colours, 1))
