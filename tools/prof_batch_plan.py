"""A/B of the batch planners: page-locked vs. heap plan arrays, number of planning workers; then one timeline."""
import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
modes = [("pinned", 4), ("heap", 4), ("pinned", 8), ("pinned", 11), ("pinned", 16), ("pinned", 6)]
for mode, workers in modes:
    if mode == "heap": os.environ["DG_NO_PINNED_PLAN"] = "1"
    else: os.environ.pop("DG_NO_PINNED_PLAN", None)
    os.environ["DG_PLAN_WORKERS"] = str(workers)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        outs = ctx.dp_diploid_batch([g] * n, 18)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert all(o["value"] == 60729 for o in outs)
    print(f"batch n={n} plan in {mode} memory, {workers} workers: " + " ".join("%.1f" % t for t in ts) + f" ms -> {n/min(ts)*1e3:.2f} samples/s", flush=True)
os.environ.pop("DG_NO_PINNED_PLAN", None)
os.environ["DG_PLAN_WORKERS"] = sys.argv[2] if len(sys.argv) > 2 else "8"
os.environ["DG_TIMING"] = "1"
ctx.dp_diploid_batch([g] * n, 18)
