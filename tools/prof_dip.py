import sys, os, json
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g,_ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
p = ctx.dip_create(g, 18)
for i in range(2):
    p.run(profile=True); r = p.result()
st = p.stats(); pf = p.profile()
print(json.dumps(dict(stats=st, profile=pf, value=r['value'])))
