import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
for conc, ctas in ((0, 8), (0, 4), (0, 6), (1, 48)):
    for rep in range(2):
        t0 = time.perf_counter()
        outs = ctx.dp_diploid_batch([g] * n, 18, max_concurrent=conc, ctas_per_sample=ctas)
        dt = time.perf_counter() - t0
    assert all(o["value"] == 60729 for o in outs)
    print(f"batch n={n} max_concurrent={conc} ctas={ctas}: {dt*1e3:.1f} ms -> {n/dt:.2f} samples/s", flush=True)
