import sys, os, json
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context, LevelGraph
g,_ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
M = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if M: g = synth.truncate_levels(g, M)
ctx = Context(0)
p = ctx.dip_create(g, 18)
for i in range(2):
    p.run(profile=True); r = p.result()
st = p.stats(); pf = p.profile()
out = dict(stats=st, profile=pf, value=r['value'])
print(json.dumps(out))
