import sys, os, json
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context, LevelGraph
import oracle
g,_ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
M = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if M: g = synth.truncate_levels(g, M)
ctx = Context(0)
p = ctx.dip_create(g, 18)
for i in range(2):
    p.run(profile=True); r = p.result()
st = p.stats(); pf = p.profile()
out = dict(stats=st, profile=pf, value=r['value'])
if M and M <= 30000:
    o = oracle.dp_diploid(g.level_off, g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.colour_is_hom, 18)
    out['oracle_value'] = o['value']
print(json.dumps(out))
