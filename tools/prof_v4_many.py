"""One fused launch of S resident problems (MHC_4, R = 18), for ncu captures and A/B tests.
Usage: prof_v4_many.py S [reps] [torch] [flush] [maxconn]"""
import os, sys
opts = set(sys.argv[3:])
if "maxconn" in opts:
    os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
sys.path.insert(0, os.getcwd())
if "torch" in opts:
    import torch
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
from dipgenie_b200.cuda_api import Context, LevelGraph
g, _ = LevelGraph.from_npz('tests/golden/mhc4_chm13_dipin.npz')
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0)
probs = [ctx.dip_create(g, 18, slot=i % 1024, ctas=1) for i in range(S)]
for rep in range(reps):
    if "flush" in opts:
        flush.fill_(1); torch.cuda.synchronize()
    ms = ctx.dip_run_many(probs)
assert all(q.result()['value'] == 60729 for q in probs)
print(f"S={S} {sorted(opts)}: group {ms:.1f} ms, fused sweep {probs[0].stats()['sweep_ms']:.1f} ms", flush=True)
