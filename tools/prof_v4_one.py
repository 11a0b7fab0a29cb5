"""One sweep of the level-program engine on the first M levels of the bench graph (for ncu captures).
Usage: prof_v4_one.py [M] [R]"""
import os, sys
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
M = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R = int(sys.argv[2]) if len(sys.argv) > 2 else 18
if M:
    g = synth.truncate_levels(g, M)
ctx = Context(0)
p = ctx.dip_create(g, R)
for i in range(2):
    p.run(); r = p.result()
st = p.stats()
print("value", r["value"], "engine", st["engine"], "grid", st["grid_ctas"], "sweep_ms %.3f" % st["sweep_ms"], flush=True)
p.close()
