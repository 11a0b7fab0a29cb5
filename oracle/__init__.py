"""oracle — TEST INFRASTRUCTURE ONLY.

ctypes bindings to oracle/liboracle.so, the plain-C CPU restatement of the
reference's hot path (each C function cites the reference file:line it follows).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (dipgenie_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile oracle/*.c into oracle/liboracle.so (gcc, C99)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in sorted(os.listdir(_HERE)) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
    return _LIB


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def dp_diploid(level_off, adj_off, adj_dst, adj_w, col_off, col_val, colour_is_hom, R, want_checksums=False):
    """Oracle diploid DP on a levelized graph. Returns dict(value, s_het, p1_edges, p2_edges[, checksum, live])."""
    level_off = _c(level_off, np.int32)
    adj_off = _c(adj_off, np.int64)
    adj_dst = _c(adj_dst, np.int32)
    adj_w = _c(adj_w, np.uint8)
    col_off = _c(col_off, np.int64)
    col_val = _c(col_val, np.int32)
    colour_is_hom = _c(colour_is_hom, np.uint8)
    L = len(level_off) - 1
    val = C.c_int32(0)
    shet = C.c_int32(0)
    p1 = np.zeros(2 * (R + 2), np.int32)
    p2 = np.zeros(2 * (R + 2), np.int32)
    n1 = C.c_int32(0)
    n2 = C.c_int32(0)
    cs = np.zeros(L, np.uint64) if want_checksums else None
    lv = np.zeros(L, np.uint64) if want_checksums else None
    f = lib().dgo_dp_diploid
    f.restype = C.c_int
    rc = f(C.c_int32(L), _p(level_off, _i32p), _p(adj_off, _i64p), _p(adj_dst, _i32p), _p(adj_w, _u8p),
           _p(col_off, _i64p), _p(col_val, _i32p), _p(colour_is_hom, _u8p), C.c_int32(len(colour_is_hom)),
           C.c_int32(R), C.byref(val), C.byref(shet), _p(p1, _i32p), C.byref(n1), _p(p2, _i32p), C.byref(n2),
           _p(cs, _u64p) if want_checksums else None, _p(lv, _u64p) if want_checksums else None)
    if rc != 0:
        raise RuntimeError(f"dgo_dp_diploid failed rc={rc}")
    out = dict(value=val.value, s_het=shet.value,
               p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(), p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy())
    if want_checksums:
        out["checksum"] = cs
        out["live"] = lv
    return out


def dp_haploid(adj_off, adj_dst, adj_w, col_off, col_val, n_colours, R):
    """Oracle haploid DP on the Kahn-ordered expanded graph. Returns dict(colours_by_r, paths[list per r])."""
    adj_off = _c(adj_off, np.int64)
    adj_dst = _c(adj_dst, np.int32)
    adj_w = _c(adj_w, np.uint8)
    col_off = _c(col_off, np.int64)
    col_val = _c(col_val, np.int32)
    n = len(adj_off) - 1
    cby = np.zeros(R + 1, np.int32)
    poff = np.zeros(R + 2, np.int64)
    cap = (R + 1) * n
    pval = np.zeros(cap, np.int32)
    f = lib().dgo_dp_haploid
    f.restype = C.c_int
    rc = f(C.c_int32(n), _p(adj_off, _i64p), _p(adj_dst, _i32p), _p(adj_w, _u8p), _p(col_off, _i64p), _p(col_val, _i32p),
           C.c_int32(n_colours), C.c_int32(R), _p(cby, _i32p), _p(poff, _i64p), _p(pval, _i32p), C.c_int64(cap))
    if rc != 0:
        raise RuntimeError(f"dgo_dp_haploid failed rc={rc}")
    return dict(colours_by_r=cby, paths=[pval[poff[r]:poff[r + 1]].copy() for r in range(R + 1)])


def _take(ptr, n, dtype):
    """Copy n items from a malloc'ed C array into numpy and free it."""
    if n == 0:
        lib().dgo_free(ptr)
        return np.zeros(0, dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    lib().dgo_free(ptr)
    return a


def hash64(kmer: bytes) -> int:
    f = lib().dgo_hash64
    f.restype = C.c_uint64
    buf = (C.c_uint8 * len(kmer)).from_buffer_copy(kmer)
    return int(f(buf, C.c_int(len(kmer))))


def sketch_reads(bases, read_off, k, w, per_read=False):
    """Oracle for dg_sketch_reads: (spectrum sorted, read_count[, per-read off, val])."""
    bases = _c(bases, np.uint8)
    read_off = _c(read_off, np.uint64)
    n = len(read_off) - 1
    sp = _u64p()
    rc = _u32p()
    ns = C.c_uint64(0)
    po = _u64p()
    pv = _u64p()
    L = lib()
    L.dgo_free.argtypes = [C.c_void_p]
    rcode = L.dgo_sketch_reads(_p(bases, _u8p), _p(read_off, _u64p), C.c_uint32(n), C.c_int(k), C.c_int(w),
                               C.byref(sp), C.byref(rc), C.byref(ns), C.byref(po) if per_read else None,
                               C.byref(pv) if per_read else None)
    if rcode != 0:
        raise RuntimeError(f"dgo_sketch_reads rc={rcode}")
    spectrum = _take(sp, ns.value, np.uint64)
    count = _take(rc, ns.value, np.uint32)
    if not per_read:
        return spectrum, count
    off = _take(po, n + 1, np.uint64)
    val = _take(pv, int(off[-1]), np.uint64)
    return spectrum, count, off, val


def index_walks(seg_bases, seg_off, walk_vtx, walk_off, top_order_map, k, w, spectrum, want_all=False):
    """Oracle for dg_index_walks."""
    seg_bases = _c(seg_bases, np.uint8)
    seg_off = _c(seg_off, np.uint64)
    walk_vtx = _c(walk_vtx, np.int32)
    walk_off = _c(walk_off, np.uint64)
    top_order_map = _c(top_order_map, np.int32)
    spectrum = _c(spectrum, np.uint64)
    nw = len(walk_off) - 1
    nmin = np.zeros(nw, np.uint64)
    ho = _u64p()
    hs = _u32p()
    vo = _u64p()
    hv = _i32p()
    ao = _u64p()
    ah = _u64p()
    L = lib()
    L.dgo_free.argtypes = [C.c_void_p]
    rcode = L.dgo_index_walks(_p(seg_bases, _u8p), _p(seg_off, _u64p), C.c_uint32(len(seg_off) - 1), _p(walk_vtx, _i32p),
                              _p(walk_off, _u64p), C.c_uint32(nw), _p(top_order_map, _i32p), C.c_int(k), C.c_int(w),
                              _p(spectrum, _u64p), C.c_uint64(len(spectrum)), _p(nmin, _u64p), C.byref(ho), C.byref(hs),
                              C.byref(vo), C.byref(hv), C.byref(ao) if want_all else None, C.byref(ah) if want_all else None)
    if rcode != 0:
        raise RuntimeError(f"dgo_index_walks rc={rcode}")
    hit_off = _take(ho, nw + 1, np.uint64)
    nh = int(hit_off[-1])
    hit_sid = _take(hs, nh, np.uint32)
    hit_vtx_off = _take(vo, nh + 1, np.uint64)
    hit_vtx = _take(hv, int(hit_vtx_off[-1]), np.int32)
    out = dict(n_minimizers=nmin, hit_off=hit_off, hit_sid=hit_sid, hit_vtx_off=hit_vtx_off, hit_vtx=hit_vtx)
    if want_all:
        out["all_off"] = _take(ao, nw + 1, np.uint64)
        out["all_hash"] = _take(ah, int(out["all_off"][-1]), np.uint64)
    return out
