"""Does a large resident group (dg_dip_run_many) slow down a following dg_dp_diploid_batch?  (memory pool / stream effects)"""
import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.getcwd())
if '--torch' in sys.argv:
    import torch
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); flush.fill_(1); torch.cuda.synchronize()
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
ctx = Context(0)
def batch(tag, n=22):
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ctx.dp_diploid_batch([g] * n, 18); ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, " ".join("%.0f" % t for t in ts), flush=True)
if "--no-first" not in sys.argv: batch("fresh context:")
S, ctas = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 1)
probs = [ctx.dip_create(g, 18, slot=i, ctas=ctas) for i in range(S)]
ctx.dip_run_many(probs)
if "--no-first" not in sys.argv: batch(f"with {S} resident problems alive:")
for p in probs: p.close()
batch(f"after closing them:")
os.environ["DG_TIMING_UPLOAD"] = "1"
ctx.dp_diploid_batch([g] * 22, 18)
