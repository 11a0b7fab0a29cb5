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
