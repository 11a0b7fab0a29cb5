import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
shapes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1:]] or [(32, 4), (37, 4), (49, 3), (74, 2), (120, 1)]
for S, ctas in shapes:
    probs = [ctx.dip_create(g, 18, slot=i, ctas=ctas) for i in range(S)]
    for rep in range(2):
        ms = ctx.dip_run_many(probs)
    assert all(p.result()['value'] == 60729 for p in probs)
    sw = [p.result() and p.stats()["sweep_ms"] for p in probs]
    print(f"S={S} ctas={ctas}: group {ms:.1f} ms  per-launch sweep min {min(sw):.1f} max {max(sw):.1f}  -> {S/ms*1e3:.1f} samples/s (device)", flush=True)
    for p in probs:
        p.close()
