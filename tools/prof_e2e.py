import sys, os, time, json
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
for i in range(4):
    t0 = time.perf_counter()
    o = ctx.dp_diploid(g, 18)
    print("one-shot %d: %.1f ms value %d" % (i, (time.perf_counter() - t0) * 1e3, o["value"]), file=sys.stderr)
