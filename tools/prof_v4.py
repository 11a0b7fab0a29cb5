"""Timing of the level-program engine on the bench workload (MHC_4 + CHM13 reads, R = 18): one sample with the
default grid, then resident groups of S samples x 1 CTA in one fused launch.  Usage: prof_v4.py [S ...]"""
import json, os, sys
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
R = int(os.environ.get("PROF_R", "18"))
ctx = Context(0)
p = ctx.dip_create(g, R)
for i in range(3):
    p.run(); r = p.result()
st = p.stats()
p.run(profile=True); p.result()
import numpy as np, ctypes as C
pr = np.zeros(24, np.uint64)
ctx.lib.dg_dip_profile(C.c_void_p(ctx.h), p.h, pr.ctypes.data_as(C.c_void_p))
print("per class [levels, cycles/level]: compact %d %.0f | hand-over %d %.0f | HBM %d %.0f" % (pr[4], pr[5] / max(1, pr[4]), pr[6], pr[7] / max(1, pr[6]), pr[8], pr[9] / max(1, pr[8])), flush=True)
print("narrow loop, warp 0: set-up %.1f M, units %.1f M cycles; %d units (big %d, multi %d, copy %d, dead %d)" % (pr[18] / 1e6, pr[19] / 1e6, pr[20], pr[21], pr[22], pr[23], 0), flush=True)
print("CTA0/thread0 cycles: slot wait %.1f M, work %.1f M, barrier %.1f M over %d levels (%.0f cycles/level)" % (pr[0] / 1e6, pr[1] / 1e6, pr[2] / 1e6, pr[3], (pr[0] + pr[1] + pr[2]) / max(1, pr[3])), flush=True)
print(json.dumps(dict(single=dict(value=r['value'], engine=st['engine'], grid=st['grid_ctas'], sweep_ms=st['sweep_ms'], trace_ms=st['traceback_ms'],
                                  plan_ms=st['plan_ms'], upload_ms=st['upload_ms'], build_ms=st['build_ms'], prog_MB=st['prog_bytes'] / 1e6,
                                  code_MB=st['code_bytes'] / 1e6, device_MB=st['device_bytes'] / 1e6, n_smem=st['n_narrow'], n_wide=st['n_wide']))), flush=True)
p.close()
# the same counters for a PACKED problem (batch-slot geometry: one CTA, small tile) running alone
p = ctx.dip_create(g, R, slot=0, ctas=1)
for i in range(2):
    p.run(); p.result()
p.run(profile=True); p.result()
st = p.stats()
pr = np.zeros(24, np.uint64)
ctx.lib.dg_dip_profile(C.c_void_p(ctx.h), p.h, pr.ctypes.data_as(C.c_void_p))
print("PACKED alone: sweep %.1f ms; per class [levels, cycles/level]: compact %d %.0f | hand-over %d %.0f | HBM %d %.0f" % (st['sweep_ms'], pr[4], pr[5] / max(1, pr[4]), pr[6], pr[7] / max(1, pr[6]), pr[8], pr[9] / max(1, pr[8])), flush=True)
print("PACKED narrow loop, warp 0: set-up %.1f M, units %.1f M cycles; %d units; slot wait %.1f M, work %.1f M, barrier %.1f M over %d levels" % (pr[18] / 1e6, pr[19] / 1e6, pr[20], pr[0] / 1e6, pr[1] / 1e6, pr[2] / 1e6, pr[3]), flush=True)
p.close()
for S in [int(a) for a in sys.argv[1:]]:
    probs = [ctx.dip_create(g, R, slot=i % 1024, ctas=1) for i in range(S)]
    for rep in range(3):
        ms = ctx.dip_run_many(probs)
    res = [q.result() for q in probs]
    assert all(x['value'] == r['value'] for x in res)
    sw = probs[0].stats()["sweep_ms"]
    print(f"S={S} x 1 CTA: group {ms:.1f} ms, fused sweep {sw:.1f} ms -> {S / ms * 1e3:.1f} samples/s (device)", flush=True)
    for q in probs:
        q.close()
