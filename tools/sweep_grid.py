import sys, os, json
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
for grid in [int(x) for x in sys.argv[1:]]:
    os.environ["DG_DIP_GRID"] = str(grid)
    p = ctx.dip_create(g, 18)
    ms = []
    for i in range(3):
        p.run(); r = p.result(); ms.append(p.stats()["sweep_ms"])
    st = p.stats()
    print(grid, st["grid_ctas"], "sweep_ms", ["%.1f" % m for m in ms], "value", r["value"], flush=True)
    p.close()
