"""Timeline of the e2e batch call (DG_TIMING=1): 22 MHC_4 samples, 6 calls, alternating with one-shot calls like bench.py."""
import sys, os, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
os.environ["DG_TIMING"] = "1"
os.environ.setdefault("DG_HOST_THREADS", str(os.cpu_count()))
sys.path.insert(0, os.getcwd())
import torch
import bench
from dipgenie_b200.cuda_api import Context
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
samples, _ = bench.load_samples("mhc4_chm13", 8, seed=2)
graphs = [samples[i % 8] for i in range(22)]
ctx = Context(0)
for rep in range(5):
    flush.fill_(1); torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = ctx.dp_diploid_batch(graphs, 18)
    dt = time.perf_counter() - t0
    print(f"=== call {rep}: {dt * 1e3:.1f} ms", file=sys.stderr, flush=True)
    ctx.dp_diploid(samples[0], 18)
