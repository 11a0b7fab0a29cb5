import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
EMU_DIR = os.path.join(ROOT, "tests", "emu")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the reference binaries under oracle/_ref (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "DipGenie")) and os.path.isdir("/root/reference/test")
    skip_ref = pytest.mark.skip(reason="reference build (oracle/_ref) or /root/reference not present")
    for it in items:
        if "ref" in it.keywords and not have_ref:
            it.add_marker(skip_ref)


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(GOLD, "expected.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def _build_emu():
    so = os.path.join(EMU_DIR, "libdpemu.so")
    srcs = [os.path.join(EMU_DIR, "dp_emu.cpp"), os.path.join(ROOT, "dipgenie_b200", "csrc", "cuda", "dp_prep.cpp")]
    deps = srcs + [os.path.join(ROOT, "dipgenie_b200", "csrc", "cuda", h) for h in ("dp_prep.h", "dp_cell.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", so] + srcs)
    return so


class DpEmu:
    """CPU emulation of the CUDA sweep's kernel logic (tests/emu/dp_emu.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(_build_emu())

    def dp_diploid(self, g, R, force_pred32=False, shape=None):
        """shape = (grid, threads, tile_cells, slot_bytes, delta_max_in, trace_T, no_pack, no_long) of the emulated kernel geometry
        (0 = default; delta_max_in < 0 forces the on-the-fly mask path)."""
        L = g.n_levels
        shp = np.zeros(8, np.int32)
        if shape:
            shp[: len(shape)] = list(shape)
        counts = np.zeros(10, np.int64)
        val, sh, n1, n2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        p1 = np.zeros(2 * (R + 2), np.int32)
        p2 = np.zeros(2 * (R + 2), np.int32)
        cs = np.zeros(L, np.uint64)
        lv = np.zeros(L, np.uint64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_dp_diploid(
            C.c_int32(L), P(g.level_off), P(g.adj_off), P(g.adj_dst), P(g.adj_w), P(g.col_off), P(g.col_val),
            P(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R), C.byref(val), C.byref(sh), P(p1),
            C.byref(n1), P(p2), C.byref(n2), P(cs), P(lv), C.c_int32(1 if force_pred32 else 0), P(shp), P(counts))
        if rc != 0:
            raise RuntimeError(f"emu_dp_diploid rc={rc}")
        return dict(value=val.value, s_het=sh.value, p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(),
                    p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy(), checksum=cs, live=lv,
                    modes=dict(narrow=int(counts[0]), wide=int(counts[1]), tasks=int(counts[2]),
                               tasks_global=int(counts[3]), tasks_masks=int(counts[4]), matrices=int(counts[5]),
                               tasks_lanes=int(counts[6]), tasks_long=int(counts[7]), tasks_lanes_inplace=int(counts[8])))


    def dp_diploid_sharded(self, g, R, n_ranks, grid=4, tile_cells=0):
        """The row-sharded sweep over `n_ranks` emulated GPUs (dg_dip_create_sharded)."""
        counts = np.zeros(8, np.int64)
        val, sh, n1, n2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        p1 = np.zeros(2 * (R + 2), np.int32)
        p2 = np.zeros(2 * (R + 2), np.int32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_dp_diploid_sharded(
            C.c_int32(g.n_levels), P(g.level_off), P(g.adj_off), P(g.adj_dst), P(g.adj_w), P(g.col_off), P(g.col_val),
            P(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R), C.c_int32(n_ranks), C.c_int32(grid),
            C.c_int32(tile_cells), C.byref(val), C.byref(sh), P(p1), C.byref(n1), P(p2), C.byref(n2), P(counts))
        if rc != 0:
            raise RuntimeError(f"emu_dp_diploid_sharded rc={rc}")
        return dict(value=val.value, s_het=sh.value, p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(),
                    p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy(),
                    modes=dict(narrow=int(counts[0]), wide=int(counts[1]), wide_tasks=int(counts[2]), pushes=int(counts[3])))


@pytest.fixture(scope="session")
def dp_emu():
    return DpEmu()


def _build_emu4():
    so = os.path.join(EMU_DIR, "libdpemu4.so")
    cu = os.path.join(ROOT, "dipgenie_b200", "csrc", "cuda")
    srcs = [os.path.join(EMU_DIR, "dp_emu4.cpp"), os.path.join(cu, "dp_prep.cpp"), os.path.join(cu, "dp_plan4.cpp")]
    deps = srcs + [os.path.join(cu, h) for h in ("dp_prep.h", "dp_cell.h", "dp_prog.h", "dp_plan4.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-fwrapv", "-fPIC", "-shared", "-o", so] + srcs)
    return so


class DpEmu4:
    """CPU emulation of the level-program sweep (engine v4: dp_prog.h / dp_plan4.cpp, tests/emu/dp_emu4.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(_build_emu4())

    def dp_diploid(self, g, R, shape=None):
        """shape = (slog, kn, slot_bytes, grid, rc) of the emulated kernel geometry (0 = default)."""
        L = g.n_levels
        shp = np.zeros(8, np.int32)
        if shape:
            shp[: len(shape)] = list(shape)
        counts = np.zeros(16, np.int64)
        val, sh, n1, n2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        p1 = np.zeros(2 * (R + 2), np.int32)
        p2 = np.zeros(2 * (R + 2), np.int32)
        cs = np.zeros(L, np.uint64)
        lv = np.zeros(L, np.uint64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu4_dp_diploid(
            C.c_int32(L), P(g.level_off), P(g.adj_off), P(g.adj_dst), P(g.adj_w), P(g.col_off), P(g.col_val),
            P(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R), C.byref(val), C.byref(sh), P(p1),
            C.byref(n1), P(p2), C.byref(n2), P(cs), P(lv), P(shp), P(counts))
        if rc == -2:
            return None          # outside what the packed-key level program covers (the task-stream engine takes it)
        if rc != 0:
            raise RuntimeError(f"emu4_dp_diploid rc={rc}")
        return dict(value=val.value, s_het=sh.value, p1_edges=p1[: 2 * n1.value].reshape(-1, 2).copy(),
                    p2_edges=p2[: 2 * n2.value].reshape(-1, 2).copy(), checksum=cs, live=lv,
                    modes=dict(smem=int(counts[0]), all_ctas=int(counts[1]), compact=int(counts[2]), staged=int(counts[3]),
                               big_cells=int(counts[4]), prog_bytes=int(counts[5]), code_elems=int(counts[6]),
                               max_cand=int(counts[7]), relocations=int(counts[8]), skipped=int(counts[9]),
                               cells_written=int(counts[10]), cells_total=int(counts[11])))


    def build_program(self, g, R, shape=None) -> np.ndarray:
        """The level programs as the host builder writes them (compare with DipProblem.debug_program())."""
        shp = np.zeros(8, np.int32)
        if shape:
            shp[: len(shape)] = list(shape)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        f = self.lib.emu4_build_program
        f.restype = C.c_int64
        args = [C.c_int32(g.n_levels), P(g.level_off), P(g.adj_off), P(g.adj_dst), P(g.adj_w), P(g.col_off), P(g.col_val),
                P(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R), P(shp)]
        n = f(*args, None, C.c_int64(0))
        if n < 0:
            return None
        out = np.zeros(n, np.uint8)
        assert f(*args, P(out), C.c_int64(n)) == n
        return out


@pytest.fixture(scope="session")
def dp_emu4():
    return DpEmu4()


def oracle_dip(oracle_mod, g, R, want_checksums=True):
    return oracle_mod.dp_diploid(g.level_off, g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.colour_is_hom, R,
                                 want_checksums=want_checksums)


def assert_dip_equal(a, b, checks=True):
    assert a["value"] == b["value"]
    assert a["s_het"] == b["s_het"]
    assert np.array_equal(a["p1_edges"], b["p1_edges"])
    assert np.array_equal(a["p2_edges"], b["p2_edges"])
    if checks and "checksum" in a and "checksum" in b:
        assert np.array_equal(a["live"][1:], b["live"][1:])
        assert np.array_equal(a["checksum"][1:], b["checksum"][1:])
