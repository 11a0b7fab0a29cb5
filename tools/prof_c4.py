"""Phase counters of the level-program sweep on a config-4 panel (bench.py: config4_graph).  Usage: prof_c4.py [denom] [R]"""
import os, sys
sys.path.insert(0, os.getcwd())
import ctypes as C
import numpy as np
import bench
from dipgenie_b200.cuda_api import Context
denom = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 18
g, _ = bench.config4_graph(90, denom)
ctx = Context(0)
p = ctx.dip_create(g, R)
for i in range(2):
    p.run(); r = p.result()
st = p.stats()
p.run(profile=True); p.result()
pr = np.zeros(24, np.uint64)
ctx.lib.dg_dip_profile(C.c_void_p(ctx.h), p.h, pr.ctypes.data_as(C.c_void_p))
print("value", r["value"], "grid", st["grid_ctas"], "sweep_ms %.2f trace_ms %.2f plan_ms %.1f build_ms %.1f" % (st["sweep_ms"], st["traceback_ms"], st["plan_ms"], st["build_ms"]))
print("per class [levels, cycles/level]: compact %d %.0f | hand-over %d %.0f | HBM %d %.0f" % (pr[4], pr[5] / max(1, pr[4]), pr[6], pr[7] / max(1, pr[6]), pr[8], pr[9] / max(1, pr[8])))
print("CTA0/thread0 cycles: slot wait %.1f M, work %.1f M (of which grid wait %.1f M), barrier %.1f M over %d levels" % (pr[0] / 1e6, pr[1] / 1e6, pr[10] / 1e6, pr[2] / 1e6, pr[3]))
print("units phases (warp 0 of CTA 0), M cycles: giant %.1f big %.1f multi %.1f copy %.1f dead %.1f" % tuple(pr[11:16] / 1e6))
p.close()
