"""Per-transition statistics of the level-program plan (CPU only): how the work of a graph splits over the
transition classes of dp_sweep4.cuh.  Usage: plan_stats.py [slog kn slot_bytes grid rc]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import _build_emu4
from dipgenie_b200.cuda_api import LevelGraph
R = int(os.environ.get("PLAN_R", "18"))
if os.environ.get("PLAN_DGD"):      # a DP input dumped by the CLI (DG_DUMP_DIPIN)
    from dipgenie_b200 import dgd
    d = dgd.load(os.environ["PLAN_DGD"])
    g = LevelGraph(d["level_off"], d["adj_off"], d["adj_dst"], d["adj_w"], d["col_off"], d["col_val"], d["colour_is_hom"])
else:
    g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
shape = np.zeros(8, np.int32)
a = [int(x) for x in sys.argv[1:]]
shape[:len(a)] = a
lib = C.CDLL(_build_emu4())
lib.emu4_plan_stats.restype = C.c_int64
out = np.zeros((g.n_levels - 1, 8), np.int64)
P = lambda x: x.ctypes.data_as(C.c_void_p)
n = lib.emu4_plan_stats(C.c_int32(g.n_levels), P(g.level_off), P(g.adj_off), P(g.adj_dst), P(g.adj_w), P(g.col_off), P(g.col_val),
                        P(g.colour_is_hom), C.c_int32(len(g.colour_is_hom)), C.c_int32(R), P(shape), P(out))
print("program bytes", n)
fl, k, k2, ncopy, nmulti, ncand, nbig, ndead = out.T
PF_COMPACT, PF_SRC, PF_DST, PF_STAGED, PF_RELOC = 1, 2, 4, 8, 128
ss, ds = (fl & PF_SRC) != 0, (fl & PF_DST) != 0
work = ncopy + nmulti + ndead
cls = {"smem": ss & ds, "handover": ss ^ ds, "hbm": ~ss & ~ds}
for name, m in cls.items():
    act = m & ((work > 0) | ((fl & PF_RELOC) != 0))
    print(f"{name:9s} transitions {m.sum():7d} active {act.sum():7d} staged {(act & ((fl & PF_STAGED) != 0)).sum():7d}  cells/level copy {ncopy[act].mean():8.1f} multi {nmulti[act].mean():8.1f} cand {ncand[act].mean():9.1f} big {nbig[act].mean():6.1f} dead {ndead[act].mean():6.1f}  k2 mean {k2[act].mean():6.1f}  sums: copy {ncopy[act].sum()/1e6:.2f}M multi {nmulti[act].sum()/1e6:.2f}M cand {ncand[act].sum()/1e6:.2f}M big {nbig[act].sum()/1e3:.0f}k")
m = cls["hbm"]
print("code elements (multi cells x RL layers): %.3e" % (nmulti.sum() * ((R + 10) // 10 * 10)))
for lo, hi in ((0, 32), (32, 48), (48, 64), (64, 96), (96, 200), (200, 400), (400, 700), (700, 100000)):
    b = m & (k2 >= lo) & (k2 < hi)
    if b.sum():
        print(f"  hbm k2 in [{lo},{hi}): {b.sum():6d} levels, copy {ncopy[b].mean():8.1f} multi {nmulti[b].mean():8.1f} cand {ncand[b].mean():9.1f} big {nbig[b].mean():6.1f}")
