"""ctypes bindings to libdipgenie_cuda.so (include/dipgenie_cuda.h).

This is the device side of the hot path; there is no CPU fallback.  `load()` raises if the
library has not been built or no CUDA device can be opened.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _build

_LIB: Optional[C.CDLL] = None


class DipStats(C.Structure):
    _fields_ = [
        ("cell_updates", C.c_uint64), ("cells", C.c_uint64), ("algo_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
        ("n_levels", C.c_int32), ("n_vertices", C.c_int32), ("max_width", C.c_int32), ("max_indegree", C.c_int32),
        ("mask_words_max", C.c_int32), ("grid_ctas", C.c_int32), ("pred_bytes", C.c_int32), ("launches", C.c_int32),
        ("sweep_ms", C.c_float), ("traceback_ms", C.c_float), ("delta_ms", C.c_float), ("plan_ms", C.c_float),
        ("upload_ms", C.c_float), ("n_narrow", C.c_int32), ("n_wide", C.c_int32), ("n_tasks", C.c_int64),
        ("delta_bytes", C.c_uint64), ("prog_bytes", C.c_uint64), ("code_bytes", C.c_uint64), ("engine", C.c_int32),
        ("build_ms", C.c_float), ("cells_written", C.c_uint64), ("n_relocate", C.c_int32), ("pad_", C.c_int32),
        ("h2d_bytes", C.c_uint64),
    ]


class HapStats(C.Structure):
    _fields_ = [
        ("cell_updates", C.c_uint64), ("cells", C.c_uint64), ("algo_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
        ("n_levels", C.c_int32), ("n_vertices", C.c_int32), ("max_width", C.c_int32), ("max_indegree", C.c_int32),
        ("launches", C.c_int32), ("sweep_ms", C.c_float), ("traceback_ms", C.c_float),
    ]


class SketchStats(C.Structure):
    _fields_ = [("bases", C.c_uint64), ("minimizers", C.c_uint64), ("spectrum", C.c_uint64), ("hits", C.c_uint64),
                ("kernel_ms", C.c_float), ("launches", C.c_int32)]


DG_BATCH_MAX_EDGES = 64


class DipInput(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("level_off", C.c_void_p), ("adj_off", C.c_void_p), ("adj_dst", C.c_void_p),
                ("adj_w", C.c_void_p), ("col_off", C.c_void_p), ("col_val", C.c_void_p), ("colour_is_hom", C.c_void_p),
                ("n_colours", C.c_int32), ("R", C.c_int32)]


class DipOutput(C.Structure):
    _fields_ = [("status", C.c_int32), ("sink_value", C.c_int32), ("sink_s_het", C.c_int32), ("n_p1", C.c_int32),
                ("n_p2", C.c_int32), ("p1_edges", C.c_int32 * (2 * DG_BATCH_MAX_EDGES)),
                ("p2_edges", C.c_int32 * (2 * DG_BATCH_MAX_EDGES))]


class DipGenieCudaError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library (must have been built in-tree by `__graft_entry__.build()`)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_build.CUDA_LIB):
            raise DipGenieCudaError(
                f"{_build.CUDA_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the device path)")
        lib = C.CDLL(os.environ.get("DG_CUDA_LIB_OVERRIDE") or _build.CUDA_LIB)      # (override: A/B builds of tools/ab)
        lib.dg_create.restype = C.c_void_p
        lib.dg_create.argtypes = [C.c_int]
        lib.dg_destroy.argtypes = [C.c_void_p]
        lib.dg_last_error.restype = C.c_char_p
        lib.dg_last_error.argtypes = [C.c_void_p]
        lib.dg_free.argtypes = [C.c_void_p]
        _LIB = lib
    return _LIB


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class LevelGraph:
    """Flat levelized ExpandedGraph — the input contract of dg_dp_diploid (see include/dipgenie_cuda.h)."""
    level_off: np.ndarray     # int32 [L+1]
    adj_off: np.ndarray       # int64 [V+1]
    adj_dst: np.ndarray       # int32 [E]
    adj_w: np.ndarray         # uint8 [E]
    col_off: np.ndarray       # int64 [V+1]
    col_val: np.ndarray       # int32
    colour_is_hom: np.ndarray  # uint8 [n_colours]

    def __post_init__(self):
        self.level_off = np.ascontiguousarray(self.level_off, np.int32)
        self.adj_off = np.ascontiguousarray(self.adj_off, np.int64)
        self.adj_dst = np.ascontiguousarray(self.adj_dst, np.int32)
        self.adj_w = np.ascontiguousarray(self.adj_w, np.uint8)
        self.col_off = np.ascontiguousarray(self.col_off, np.int64)
        self.col_val = np.ascontiguousarray(self.col_val, np.int32)
        self.colour_is_hom = np.ascontiguousarray(self.colour_is_hom, np.uint8)

    @property
    def n_levels(self) -> int:
        return len(self.level_off) - 1

    @property
    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.level_off, self.adj_off, self.adj_dst, self.adj_w, self.col_off,
                                      self.col_val, self.colour_is_hom))

    @staticmethod
    def from_dgd(d, prefix: str = "dip_in.") -> "LevelGraph":
        return LevelGraph(d[prefix + "lvl_vtx.off"], d[prefix + "adj.off"], d[prefix + "adj.dst"], d[prefix + "adj.w"],
                          d[prefix + "color.off"], d[prefix + "color.val"], d[prefix + "color_homo"])

    def to_npz(self, path: str, **extra) -> None:
        outdeg = np.diff(self.adj_off)
        ncol = np.diff(self.col_off)
        np.savez_compressed(
            path, level_off=self.level_off,
            outdeg=outdeg.astype(np.uint16 if outdeg.max(initial=0) < 65536 else np.int32),
            adj_dst_delta=np.diff(self.adj_dst, prepend=0).astype(np.int32), adj_w=np.packbits(self.adj_w),
            n_edges=np.int64(len(self.adj_w)),
            ncol=ncol.astype(np.uint16 if ncol.max(initial=0) < 65536 else np.int32), col_val=self.col_val,
            colour_is_hom=np.packbits(self.colour_is_hom), n_colours=np.int64(len(self.colour_is_hom)), **extra)

    @staticmethod
    def from_npz(path: str):
        z = np.load(path)
        ne = int(z["n_edges"])
        g = LevelGraph(
            z["level_off"], np.concatenate([[0], np.cumsum(z["outdeg"].astype(np.int64))]),
            np.cumsum(z["adj_dst_delta"].astype(np.int64)).astype(np.int32), np.unpackbits(z["adj_w"])[:ne],
            np.concatenate([[0], np.cumsum(z["ncol"].astype(np.int64))]), z["col_val"],
            np.unpackbits(z["colour_is_hom"])[: int(z["n_colours"])])
        return g, z


@dataclass
class HapGraph:
    """Flat Kahn-ordered ExpandedGraph — the input contract of dg_dp_haploid (see include/dipgenie_cuda.h)."""
    adj_off: np.ndarray       # int64 [n+1]
    adj_dst: np.ndarray       # int32 [E]
    adj_w: np.ndarray         # uint8 [E]
    col_off: np.ndarray       # int64 [n+1]
    col_val: np.ndarray       # int32
    n_colours: int = 0

    def __post_init__(self):
        self.adj_off = np.ascontiguousarray(self.adj_off, np.int64)
        self.adj_dst = np.ascontiguousarray(self.adj_dst, np.int32)
        self.adj_w = np.ascontiguousarray(self.adj_w, np.uint8)
        self.col_off = np.ascontiguousarray(self.col_off, np.int64)
        self.col_val = np.ascontiguousarray(self.col_val, np.int32)
        if not self.n_colours:
            self.n_colours = int(self.col_val.max(initial=-1)) + 1

    @property
    def n(self) -> int:
        return len(self.adj_off) - 1

    @property
    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.adj_off, self.adj_dst, self.adj_w, self.col_off, self.col_val))

    @staticmethod
    def from_dgd(d, prefix: str = "hap_in.") -> "HapGraph":
        return HapGraph(d[prefix + "adj.off"], d[prefix + "adj.dst"], d[prefix + "adj.w"], d[prefix + "color.off"],
                        d[prefix + "color.val"])

    @staticmethod
    def from_npz(path: str) -> "HapGraph":
        z = np.load(path)
        ne = int(z["n_edges"])
        return HapGraph(np.concatenate([[0], np.cumsum(z["outdeg"].astype(np.int64))]),
                        np.cumsum(z["adj_dst_delta"].astype(np.int64)).astype(np.int32), np.unpackbits(z["adj_w"])[:ne],
                        np.concatenate([[0], np.cumsum(z["ncol"].astype(np.int64))]), z["col_val"])


class Context:
    """One dg_ctx (one GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = self.lib.dg_create(device)
        if not self.h:
            raise DipGenieCudaError(f"dg_create({device}) failed: no usable CUDA device (no CPU fallback exists)")

    def close(self):
        if self.h:
            self.lib.dg_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc: int, what: str):
        if rc != 0:
            raise DipGenieCudaError(f"{what} failed ({rc}): {self.lib.dg_last_error(self.h).decode()}")

    def device_info(self):
        sm = C.c_int(0)
        fr = C.c_size_t(0)
        tot = C.c_size_t(0)
        self.check(self.lib.dg_device_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot)), "dg_device_info")
        return dict(sm_count=sm.value, free_bytes=fr.value, total_bytes=tot.value)

    # ---- diploid DP ------------------------------------------------------------------------
    def dp_diploid(self, g: LevelGraph, R: int):
        """One-shot host-buffer call (dg_dp_diploid): H2D, sweep, traceback, D2H."""
        val = C.c_int32(0)
        shet = C.c_int32(0)
        n1 = C.c_int32(0)
        n2 = C.c_int32(0)
        p1 = np.zeros(2 * (R + 2), np.int32)
        p2 = np.zeros(2 * (R + 2), np.int32)
        rc = self.lib.dg_dp_diploid(
            C.c_void_p(self.h), C.c_int32(g.n_levels), _ptr(g.level_off), _ptr(g.adj_off), _ptr(g.adj_dst), _ptr(g.adj_w),
            _ptr(g.col_off), _ptr(g.col_val), _ptr(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R),
            C.byref(val), C.byref(shet), _ptr(p1), C.byref(n1), _ptr(p2), C.byref(n2))
        self.check(rc, "dg_dp_diploid")
        return dict(value=val.value, s_het=shet.value, p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(),
                    p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy())

    def dp_diploid_batch(self, graphs, R, max_concurrent: int = 0, ctas_per_sample: int = 0):
        """dg_dp_diploid_batch: independent samples resident on the GPU together; R is an int or one per sample."""
        n = len(graphs)
        Rs = [int(R)] * n if np.isscalar(R) else [int(r) for r in R]
        ins = (DipInput * max(n, 1))()
        outs = (DipOutput * max(n, 1))()
        for i, g in enumerate(graphs):
            ins[i] = DipInput(g.n_levels, g.level_off.ctypes.data, g.adj_off.ctypes.data, g.adj_dst.ctypes.data,
                              g.adj_w.ctypes.data, g.col_off.ctypes.data, g.col_val.ctypes.data, g.colour_is_hom.ctypes.data,
                              len(g.colour_is_hom), Rs[i])
        rc = self.lib.dg_dp_diploid_batch(C.c_void_p(self.h), C.c_int32(n), ins, outs, C.c_int32(max_concurrent),
                                          C.c_int32(ctas_per_sample))
        self.check(rc, "dg_dp_diploid_batch")
        res = []
        for i in range(n):
            o = outs[i]
            res.append(dict(value=o.sink_value, s_het=o.sink_s_het,
                            p1_edges=np.array(o.p1_edges[: 2 * o.n_p1], np.int32).reshape(-1, 2),
                            p2_edges=np.array(o.p2_edges[: 2 * o.n_p2], np.int32).reshape(-1, 2)))
        return res

    def release_cached_memory(self):
        """dg_release_cached_memory: return the pool's cached device blocks (after closing many resident problems)."""
        self.check(self.lib.dg_release_cached_memory(C.c_void_p(self.h)), "dg_release_cached_memory")

    def dip_create(self, g: LevelGraph, R: int, slot: Optional[int] = None, ctas: int = 0) -> "DipProblem":
        return DipProblem(self, g, R, slot, ctas)

    def dip_create_sharded(self, g: LevelGraph, R: int, rank: int, world: int, ctas: int = 0) -> "DipProblem":
        """One rank's part of a diploid DP whose wide transitions are row-split over `world` GPUs (shard.RowShardedDip)."""
        return DipProblem(self, g, R, None, ctas, shard=(rank, world))

    def dip_sharded_in_process(self, g: LevelGraph, R: int, world: int, ctas: int = 4) -> list:
        """All `world` ranks of a row-sharded problem as sibling problems on this one GPU (dg_dip_attach_in_process)."""
        probs = [self.dip_create_sharded(g, R, q, world, ctas) for q in range(world)]
        arr = (C.c_void_p * world)(*[p.h for p in probs])
        self.check(self.lib.dg_dip_attach_in_process(C.c_void_p(self.h), arr, C.c_int32(world)), "dg_dip_attach_in_process")
        return probs

    @staticmethod
    def run_sharded_siblings(probs) -> list:
        """arm all, launch all (asynchronous, own streams), collect all."""
        for p in probs:
            p.shard_arm()
        for p in probs:
            p.run()
        return [p.result() for p in probs]

    def dip_run_many(self, problems) -> float:
        """dg_dip_run_many: run resident problems (distinct slots) together; returns the group's device ms."""
        arr = (C.c_void_p * max(len(problems), 1))(*[p.h for p in problems])
        ms = C.c_float(0)
        self.check(self.lib.dg_dip_run_many(C.c_void_p(self.h), arr, C.c_int32(len(problems)), C.byref(ms)), "dg_dip_run_many")
        return ms.value

    # ---- sketch / spectrum / join ------------------------------------------------------------
    def _take(self, ptr, n, dtype):
        a = np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),))[: int(n)].astype(dtype, copy=True)
        self.lib.dg_free(ptr)
        return a

    def sketch_stats(self) -> dict:
        s = SketchStats()
        self.check(self.lib.dg_sketch_last_stats(C.c_void_p(self.h), C.byref(s)), "dg_sketch_last_stats")
        return {k: getattr(s, k) for k, _ in SketchStats._fields_}

    def sketch_minimizers(self, bases, seq_off, k: int, w: int):
        """dg_sketch_minimizers: (count per sequence, hashes, starts) in sequence order."""
        bases = np.ascontiguousarray(bases, np.uint8)
        seq_off = np.ascontiguousarray(seq_off, np.uint64)
        n = len(seq_off) - 1
        cnt = np.zeros(max(n, 1), np.uint64)
        hp = C.POINTER(C.c_uint64)()
        sp = C.POINTER(C.c_uint64)()
        rc = self.lib.dg_sketch_minimizers(C.c_void_p(self.h), _ptr(bases), _ptr(seq_off), C.c_uint32(n), C.c_int(k), C.c_int(w),
                                           _ptr(cnt), C.byref(hp), C.byref(sp))
        self.check(rc, "dg_sketch_minimizers")
        tot = int(cnt[:n].sum())
        return cnt[:n], self._take(hp, tot, np.uint64), self._take(sp, tot, np.uint64)

    def sketch_reads(self, bases, read_off, k: int, w: int):
        """dg_sketch_reads: (spectrum ascending, read_count)."""
        bases = np.ascontiguousarray(bases, np.uint8)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        sp = C.POINTER(C.c_uint64)()
        rc_ = C.POINTER(C.c_uint32)()
        ns = C.c_uint64(0)
        rc = self.lib.dg_sketch_reads(C.c_void_p(self.h), _ptr(bases), _ptr(read_off), C.c_uint32(len(read_off) - 1), C.c_int(k),
                                      C.c_int(w), C.byref(sp), C.byref(rc_), C.byref(ns))
        self.check(rc, "dg_sketch_reads")
        return self._take(sp, ns.value, np.uint64), self._take(rc_, ns.value, np.uint32)

    def index_walks(self, seg_bases, seg_off, walk_vtx, walk_off, top_order_map, k: int, w: int, spectrum):
        """dg_index_walks: dict(n_minimizers, hit_off, hit_sid, hit_vtx_off, hit_vtx)."""
        seg_bases = np.ascontiguousarray(seg_bases, np.uint8)
        seg_off = np.ascontiguousarray(seg_off, np.uint64)
        walk_vtx = np.ascontiguousarray(walk_vtx, np.int32)
        walk_off = np.ascontiguousarray(walk_off, np.uint64)
        top_order_map = np.ascontiguousarray(top_order_map, np.int32)
        spectrum = np.ascontiguousarray(spectrum, np.uint64)
        nw = len(walk_off) - 1
        nmin = np.zeros(max(nw, 1), np.uint64)
        ho = C.POINTER(C.c_uint64)()
        hs = C.POINTER(C.c_uint32)()
        vo = C.POINTER(C.c_uint64)()
        hv = C.POINTER(C.c_int32)()
        rc = self.lib.dg_index_walks(C.c_void_p(self.h), _ptr(seg_bases), _ptr(seg_off), C.c_uint32(len(seg_off) - 1),
                                     _ptr(walk_vtx), _ptr(walk_off), C.c_uint32(nw), _ptr(top_order_map), C.c_int(k), C.c_int(w),
                                     _ptr(spectrum), C.c_uint64(len(spectrum)), _ptr(nmin), C.byref(ho), C.byref(hs),
                                     C.byref(vo), C.byref(hv))
        self.check(rc, "dg_index_walks")
        hit_off = self._take(ho, nw + 1, np.uint64)
        nh = int(hit_off[-1])
        hit_sid = self._take(hs, nh, np.uint32)
        hit_vtx_off = self._take(vo, nh + 1, np.uint64)
        hit_vtx = self._take(hv, int(hit_vtx_off[-1]), np.int32)
        return dict(n_minimizers=nmin[:nw], hit_off=hit_off, hit_sid=hit_sid, hit_vtx_off=hit_vtx_off, hit_vtx=hit_vtx)

    # ---- haploid DP ------------------------------------------------------------------------
    def dp_haploid(self, g: "HapGraph", R: int):
        """One-shot host-buffer call (dg_dp_haploid)."""
        cby = np.zeros(R + 1, np.int32)
        poff = np.zeros(R + 2, np.int64)
        buf = C.POINTER(C.c_int32)()
        rc = self.lib.dg_dp_haploid(C.c_void_p(self.h), C.c_int32(g.n), _ptr(g.adj_off), _ptr(g.adj_dst), _ptr(g.adj_w),
                                    _ptr(g.col_off), _ptr(g.col_val), C.c_int32(g.n_colours), C.c_int32(R), _ptr(cby),
                                    _ptr(poff), C.byref(buf))
        self.check(rc, "dg_dp_haploid")
        tot = int(poff[-1])
        flat = np.ctypeslib.as_array(buf, shape=(max(tot, 1),))[:tot].copy()
        self.lib.dg_free(buf)
        return dict(colours_by_r=cby, paths=[flat[poff[r]:poff[r + 1]] for r in range(R + 1)])

    def hap_create(self, g: "HapGraph", R: int) -> "HapProblem":
        return HapProblem(self, g, R)


class HapProblem:
    """A haploid DP problem resident in HBM (dg_hap_*)."""

    def __init__(self, ctx: "Context", g: "HapGraph", R: int):
        self.ctx = ctx
        self.R = R
        self.n = g.n
        h = C.c_void_p(None)
        rc = ctx.lib.dg_hap_create(C.c_void_p(ctx.h), C.c_int32(g.n), _ptr(g.adj_off), _ptr(g.adj_dst), _ptr(g.adj_w),
                                   _ptr(g.col_off), _ptr(g.col_val), C.c_int32(g.n_colours), C.c_int32(R), C.byref(h))
        ctx.check(rc, "dg_hap_create")
        self.h = h

    def run(self):
        self.ctx.check(self.ctx.lib.dg_hap_run(C.c_void_p(self.ctx.h), self.h), "dg_hap_run")

    def result(self):
        cby = np.zeros(self.R + 1, np.int32)
        plen = np.zeros(self.R + 1, np.int32)
        self.ctx.check(self.ctx.lib.dg_hap_result(C.c_void_p(self.ctx.h), self.h, _ptr(cby), _ptr(plen)), "dg_hap_result")
        return dict(colours_by_r=cby, path_len=plen)

    def path(self, r: int, cap: int):
        buf = np.zeros(max(cap, 1), np.int32)
        n = C.c_int32(0)
        self.ctx.check(self.ctx.lib.dg_hap_path(C.c_void_p(self.ctx.h), self.h, C.c_int32(r), _ptr(buf), C.c_int32(cap),
                                                C.byref(n)), "dg_hap_path")
        return buf[: n.value].copy()

    def stats(self) -> dict:
        s = HapStats()
        self.ctx.check(self.ctx.lib.dg_hap_stats(C.c_void_p(self.ctx.h), self.h, C.byref(s)), "dg_hap_stats")
        return {k: getattr(s, k) for k, _ in HapStats._fields_}

    def close(self):
        if self.h:
            self.ctx.lib.dg_hap_destroy(C.c_void_p(self.ctx.h), self.h)
            self.h = None


class DipProblem:
    """A diploid DP problem resident in HBM (dg_dip_*)."""

    IPC_HANDLE_BYTES = 64          # include/dipgenie_cuda.h: DG_IPC_HANDLE_BYTES

    def __init__(self, ctx: Context, g: LevelGraph, R: int, slot: Optional[int] = None, ctas: int = 0,
                 shard: Optional[tuple] = None):
        self.ctx = ctx
        self.R = R
        self.L = g.n_levels
        self.shard = shard
        h = C.c_void_p(None)
        args = [C.c_void_p(ctx.h), C.c_int32(g.n_levels), _ptr(g.level_off), _ptr(g.adj_off), _ptr(g.adj_dst), _ptr(g.adj_w),
                _ptr(g.col_off), _ptr(g.col_val), _ptr(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R)]
        if shard is not None:      # (rank, world): row-sharded over the GPUs of the node (dg_dip_create_sharded)
            ctx.check(ctx.lib.dg_dip_create_sharded(*args, C.c_int32(shard[0]), C.c_int32(shard[1]), C.c_int32(ctas), C.byref(h)),
                      "dg_dip_create_sharded")
        elif slot is None:
            ctx.check(ctx.lib.dg_dip_create(*args, C.byref(h)), "dg_dip_create")
        else:
            ctx.check(ctx.lib.dg_dip_create_slot(*args, C.c_int32(slot), C.c_int32(ctas), C.byref(h)), "dg_dip_create_slot")
        self.h = h

    def ipc_export(self) -> bytes:
        """dg_dip_ipc_export: the 4 CUDA IPC handles (layer tiles, predecessor codes, counters) the peers map."""
        buf = (C.c_uint8 * (4 * self.IPC_HANDLE_BYTES))()
        self.ctx.check(self.ctx.lib.dg_dip_ipc_export(C.c_void_p(self.ctx.h), self.h, buf), "dg_dip_ipc_export")
        return bytes(buf)

    def ipc_attach(self, all_handles: Sequence[bytes]):
        """dg_dip_ipc_attach: `all_handles[q]` = rank q's ipc_export()."""
        blob = b"".join(all_handles)
        if len(blob) != self.shard[1] * 4 * self.IPC_HANDLE_BYTES:
            raise ValueError("ipc_attach: need one 256-byte handle block per rank")
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.ctx.check(self.ctx.lib.dg_dip_ipc_attach(C.c_void_p(self.ctx.h), self.h, buf), "dg_dip_ipc_attach")

    def shard_arm(self):
        """dg_dip_shard_arm: reset counters and level 0; every rank must do this and pass a host barrier before run()."""
        self.ctx.check(self.ctx.lib.dg_dip_shard_arm(C.c_void_p(self.ctx.h), self.h), "dg_dip_shard_arm")

    def run(self, checksums: bool = False, profile: bool = False):
        flags = (1 if checksums else 0) | (2 if profile else 0)
        self.ctx.check(self.ctx.lib.dg_dip_run(C.c_void_p(self.ctx.h), self.h, C.c_uint32(flags)), "dg_dip_run")

    def profile(self) -> dict:
        out = np.zeros(24, np.uint64)
        self.ctx.check(self.ctx.lib.dg_dip_profile(C.c_void_p(self.ctx.h), self.h, _ptr(out)), "dg_dip_profile")
        names = ["tasks", "slot_wait", "grid_wait", "cell_loop", "barrier_arrive"]
        d = {m: {n: int(out[i * 6 + j]) for j, n in enumerate(names)} for i, m in enumerate(["smem_layers", "hbm_layers"])}
        d["lane_form_warp0"] = {n: int(out[12 + j]) for j, n in enumerate(["items", "setup", "loop", "reduce", "store", "iters"])}
        return d

    def debug_program(self) -> np.ndarray:
        """dg_dip_debug_program: the level programs the device builder wrote (engine 4; empty otherwise)."""
        n = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.dg_dip_debug_program(C.c_void_p(self.ctx.h), self.h, None, C.c_uint64(0), C.byref(n)),
                       "dg_dip_debug_program")
        out = np.zeros(int(n.value), np.uint8)
        if n.value:
            self.ctx.check(self.ctx.lib.dg_dip_debug_program(C.c_void_p(self.ctx.h), self.h, _ptr(out), C.c_uint64(n.value),
                                                             C.byref(n)), "dg_dip_debug_program")
        return out

    def result(self):
        R = self.R
        val = C.c_int32(0)
        shet = C.c_int32(0)
        n1 = C.c_int32(0)
        n2 = C.c_int32(0)
        p1 = np.zeros(2 * (R + 2), np.int32)
        p2 = np.zeros(2 * (R + 2), np.int32)
        rc = self.ctx.lib.dg_dip_result(C.c_void_p(self.ctx.h), self.h, C.byref(val), C.byref(shet), _ptr(p1), C.byref(n1),
                                        _ptr(p2), C.byref(n2))
        self.ctx.check(rc, "dg_dip_result")
        return dict(value=val.value, s_het=shet.value, p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(),
                    p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy())

    def checksums(self):
        cs = np.zeros(self.L, np.uint64)
        lv = np.zeros(self.L, np.uint64)
        self.ctx.check(self.ctx.lib.dg_dip_checksums(C.c_void_p(self.ctx.h), self.h, _ptr(cs), _ptr(lv)), "dg_dip_checksums")
        return cs, lv

    def stats(self) -> dict:
        s = DipStats()
        self.ctx.check(self.ctx.lib.dg_dip_stats(C.c_void_p(self.ctx.h), self.h, C.byref(s)), "dg_dip_stats")
        return {k: getattr(s, k) for k, _ in DipStats._fields_}

    def close(self):
        if self.h:
            self.ctx.lib.dg_dip_destroy(C.c_void_p(self.ctx.h), self.h)
            self.h = None
