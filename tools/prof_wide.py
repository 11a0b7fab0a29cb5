import sys, os, json, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context
ctx = Context(0)
shapes = ((16, 200, 18), (48, 60, 18), (90, 24, 18), (320, 12, 18), (640, 8, 18)) if len(sys.argv) < 2 else [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]
for H, nb, R in shapes:
    g = synth.lane_panel_graph(90, n_lanes=H, n_blocks=nb, rec_per_block=max(2, H // 16), p_colour=0.08, n_colours=1 << 15)
    p = ctx.dip_create(g, R)
    for i in range(2):
        p.run(); r = p.result()
    st = p.stats()
    print(json.dumps(dict(H=H, levels=st["n_levels"], max_width=st["max_width"], cell_updates=st["cell_updates"], sweep_ms=st["sweep_ms"],
                          trace_ms=st["traceback_ms"], delta_ms=st["delta_ms"], grid=st["grid_ctas"], n_narrow=st["n_narrow"], n_wide=st["n_wide"],
                          gcu_per_s=st["cell_updates"] / st["sweep_ms"] / 1e6, algo_GBs=st["algo_bytes"] / st["sweep_ms"] / 1e6,
                          value=r["value"])), flush=True)      # parity of these shapes: tests/test_dp_diploid_gpu.py
    p.close()
