"""Row-sharded diploid DP over the GPUs of one node, one process per GPU (launch with torchrun):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/perf/run_row_sharded.py [--check]
Every rank runs the same problems through dipgenie_b200.shard.RowShardedDip (rows and barrier arrivals exchanged inside
the sweep kernel over NVLink peer mappings) and compares with its own single-GPU run of the same problem; rank 0 prints
the sweep times of both."""
import argparse, json, os, sys
sys.path.insert(0, os.getcwd())
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("NCCL_DEBUG", "WARN")
import numpy as np
import torch
import torch.distributed as dist
from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context, LevelGraph
from dipgenie_b200.shard import RowShardedDip

ap = argparse.ArgumentParser()
ap.add_argument("--check", action="store_true", help="small problems only")
ap.add_argument("--lanes", type=int, nargs="*", default=[48, 90, 160])
ap.add_argument("--R", type=int, default=18)
ap.add_argument("--profile", action="store_true", help="phase cycle counters of CTA 0 / thread 0 of every rank")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)


def same(x, y):
    return (x["value"] == y["value"] and x["s_het"] == y["s_het"] and np.array_equal(x["p1_edges"], y["p1_edges"])
            and np.array_equal(x["p2_edges"], y["p2_edges"]))


problems = []
lanes = [24, 40] if a.check else a.lanes
for H in lanes:
    problems.append((f"lane panel H={H}", synth.lane_panel_graph(90 + H, n_lanes=H, n_blocks=12 if a.check else 24, rec_per_block=2,
                                                                p_colour=0.1, n_colours=2048), 4 if a.check else a.R))
problems.append(("MHC_4 + CHM13 reads", LevelGraph.from_npz("tests/golden/mhc4_chm13_dipin.npz")[0], a.R))
rows, ok = [], True
for name, g, R in problems:
    single = ctx.dip_create(g, R)
    t1 = []
    for _ in range(3):
        single.run(); ref = single.result(); t1.append(single.stats()["sweep_ms"])
    single.close()
    prob = RowShardedDip(ctx, g, R, dist)
    tn = []
    for _ in range(3):
        out = prob.run(); tn.append(prob.stats()["sweep_ms"])
    st = prob.stats()
    if a.profile:
        prob.run(profile=True)
        print(f"rank {rank} {name}: {json.dumps(prob.problem.profile()['hbm_layers'])}", flush=True)
    prob.close()
    good = same(ref, out)
    flags = [None] * world
    dist.all_gather_object(flags, (good, min(tn)))
    ok &= all(f[0] for f in flags)
    rows.append(dict(problem=name, R=R, levels=int(g.n_levels), max_width=int(st["max_width"]), value=int(out["value"]), world=world,
                     ctas_per_rank=int(st["grid_ctas"]), n_narrow=int(st["n_narrow"]), n_wide=int(st["n_wide"]),
                     sweep_ms_single_gpu=round(min(t1), 3), sweep_ms_sharded_max_over_ranks=round(max(f[1] for f in flags), 3),
                     identical_on_every_rank=all(f[0] for f in flags)))
if rank == 0:
    print(json.dumps(dict(world=world, results=rows)))
    print("row-sharded OK" if ok else "row-sharded MISMATCH")
dist.barrier()
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
