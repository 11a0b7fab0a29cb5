"""e2e batch call (dg_dp_diploid_batch, host buffers) on 22 MHC_4 samples: planner workers x CTAs per sample.  DG_TIMING=1 for the timeline."""
import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
import bench
from dipgenie_b200.cuda_api import Context
samples, _ = bench.load_samples("mhc4_chm13", 8, seed=2)
graphs = [samples[i % 8] for i in range(22)]
ctx = Context(0)
for workers, ctas in ((0, 0), (11, 0), (16, 0), (8, 6), (11, 6), (6, 0)):
    if workers:
        os.environ["DG_PLAN_WORKERS"] = str(workers)
    else:
        os.environ.pop("DG_PLAN_WORKERS", None)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        outs = ctx.dp_diploid_batch(graphs, 18, ctas_per_sample=ctas)
        ts.append(time.perf_counter() - t0)
    print(f"workers={workers or 'default'} ctas={ctas or 'default'}: {[round(t * 1e3, 1) for t in ts]} ms", flush=True)
