#!/usr/bin/env python
"""bench.py — DipGenie hot path on B200: diploid recombination-constrained DP (sweep + traceback).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU solver on the host cores

A step = one full pass of the DP (sweep + traceback) over one batch of samples: S samples per GPU (default 256: two
sweep CTAs per SM, one per sample), every sample a levelized graph of BASELINE config 2's shape (MHC_4 panel, -p2 -R18;
the samples differ in their read-derived colours, see load_samples), resident in HBM (level programs + predecessor
codes, about 0.53 GB each), all swept by ONE fused launch and traced back by one launch set.  The DP of one H=5 sample
is a chain of ~10^5 dependent level transitions that keeps one SM partly busy, so samples side by side is how the path
fills a B200 (the reference's own batch use: data/run_DipGenie_batch.sh).  With N ranks every rank owns its own S
samples (no collective) -> weak scaling; value = cell-updates of all samples / max-over-ranks device time.
e2e = the same through dg_dp_diploid_batch with host buffers (22 samples per GPU: planning, H2D, program build, sweep,
traceback, D2H in the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

# batch slots are concurrent persistent kernels: one hardware work queue each (see dg_create); must be set
# before the first CUDA call of the process
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# rank 0 must print exactly one JSON line: keep NCCL's version banner off stdout
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
REF_PLAIN = os.path.join(ROOT, "oracle", "_ref", "ref_driver_plain")


# dram__bytes_read.sum + dram__bytes_write.sum of the fused sweep launch PER SAMPLE, with the ncu capture it comes from
# (a DRAM counter cannot be read from inside the timed run; the capture is of the same command and configuration).
# None where no capture exists for the configuration.
NCU_TRAFFIC = {
    # (workload, R, samples per GPU): (bytes per sample, capture)
}
try:
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as _f:
        for _e in json.load(_f):
            NCU_TRAFFIC[(_e["workload"], _e["R"], _e["samples_per_gpu"])] = (_e["dram_bytes_per_sample"], _e["capture"])
except OSError:
    pass


def load_workload(name: str):
    from dipgenie_b200 import synth
    from dipgenie_b200.cuda_api import LevelGraph
    if name == "mhc4_chm13":
        g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
        desc = "diploid -p2 DP on test/MHC_4.gfa.gz (5 walks) + test/CHM13_reads.fq.gz anchors; levelized graph from the reference front end"
        return g, desc
    m = re.fullmatch(r"c4_h(\d+)(?:_s(\d+))?", name)
    if m:
        return config4_graph(int(m.group(1)), int(m.group(2) or 1))
    m = re.fullmatch(r"lanes(\d+)x(\d+)", name)
    if m:
        H, nb = int(m.group(1)), int(m.group(2))
        g = synth.lane_panel_graph(90, n_lanes=H, n_blocks=nb, rec_per_block=max(2, H // 16), p_colour=0.08, n_colours=1 << 15)
        return g, f"synthetic lane-panel model, {H} haplotype lanes, {nb} blocks (seed 90)"
    raise SystemExit(f"unknown workload {name}")


def config4_graph(walks: int, denom: int):
    """BASELINE config 4: the seeded synthetic panel (seed 90: 5 Mbp / denom backbone, 33 000 / denom sites, 12
    founders, `walks` mosaic walks) and its 30x read set, pushed through this repo's own front end (the `dipgenie` CLI:
    sketch + join on the GPU, anchors, classifier fit, expansion, levelization on the host) up to the DP's input, which the
    CLI dumps (DG_DUMP_DIPIN).  Cached under /tmp for the R sweep."""
    from dipgenie_b200 import _build, dgd
    from dipgenie_b200.cuda_api import LevelGraph
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_config4
    work = os.path.join(tempfile.gettempdir(), "dg_c4_h%d_s%d" % (walks, denom))
    dump = os.path.join(work, "dipin.dgd")
    if not os.path.exists(dump):
        gfa, fa = make_config4.make(work, scale=1.0 / denom, walks=walks)
        env = dict(os.environ, DG_DUMP_DIPIN=dump + ".tmp", DG_DUMP_ONLY="1")
        t0 = time.perf_counter()
        p = subprocess.run([_build.CLI_BIN, "-g", gfa, "-r", fa, "-o", os.path.join(work, "out.fa"), "-t", str(os.cpu_count() or 8), "-p2", "-R18"],
                           capture_output=True, text=True, env=env, timeout=3000)
        if p.returncode != 0:
            raise SystemExit("bench.py: the front end failed on the config-4 panel: " + p.stderr[-1500:])
        os.replace(dump + ".tmp", dump)
        print("bench: config-4 front end (sketch, join, anchors, fit, expansion, levelization): %.1f s" % (time.perf_counter() - t0), file=sys.stderr)
    d = dgd.load(dump)
    g = LevelGraph(d["level_off"], d["adj_off"], d["adj_dst"], d["adj_w"], d["col_off"], d["col_val"], d["colour_is_hom"])
    desc = ("synthetic config-4 panel (seed 90): %d walks, backbone %d bp, %d sites, 30x reads of a diploid mosaic target, "
            "levelized by this repo's front end" % (walks, 5_000_000 // denom, 33_000 // denom))
    return g, desc


def wide_panel_arm(args):
    """One wide-panel sample on the whole GPU (config 4): a step = one pass of the DP (program sweep + traceback) of the
    resident problem; e2e = dg_dp_diploid with host buffers (planning, H2D, program build, sweep, traceback, D2H)."""
    import torch
    from dipgenie_b200.cuda_api import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the device path has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py: the wide-panel workload is a single-GPU workload (replicas only)")
    torch.cuda.set_device(0)
    g, desc = load_workload(args.workload)
    ctx = Context(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e2e_t = []
    for i in range(max(1, args.e2e_steps) + 1):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        o1 = ctx.dp_diploid(g, args.R)
        if i > 0:
            e2e_t.append(time.perf_counter() - t0)
    ctx.release_cached_memory()
    p = ctx.dip_create(g, args.R)
    ms, sweep, trace = [], [], []
    with ClockSampler(0) as clk:
        for i in range(args.warmup + args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            p.run()
            out = p.result()
            st = p.stats()
            if i >= args.warmup:
                ms.append(st["sweep_ms"] + st["traceback_ms"] + st["delta_ms"]); sweep.append(st["sweep_ms"]); trace.append(st["traceback_ms"])
    assert out["value"] == o1["value"] and np.array_equal(out["p1_edges"], o1["p1_edges"]) and np.array_equal(out["p2_edges"], o1["p2_edges"])
    peak, peak_src = peaks()
    dev_ms, sweep_ms = float(np.mean(ms)), float(np.mean(sweep))
    U, algo = float(st["cell_updates"]), float(st["algo_bytes"])
    achieved = algo / (sweep_ms * 1e-3) / 1e9
    e2e_s = float(np.mean(e2e_t))
    line = {
        "metric": "dp_cell_updates_per_sec", "value": U / (dev_ms * 1e-3), "unit": "cell-updates/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic (seeded generator, SURVEY 8d config 4)",
        "config": {"workload": args.workload, "description": desc, "R": args.R, "ploidy": 2, "levels": st["n_levels"], "vertices": st["n_vertices"],
                   "max_width": st["max_width"], "max_indegree": st["max_indegree"], "cell_updates_per_sample": U, "dest_cells_per_sample": st["cells"],
                   "samples_per_step": 1, "ctas_per_sample": st["grid_ctas"], "engine": st["engine"],
                   "device_bytes": st["device_bytes"], "program_bytes": st["prog_bytes"], "code_bytes": st["code_bytes"],
                   "l2": "256 MiB device buffer rewritten between timed iterations",
                   "timing": "CUDA events on the problem's stream (sweep + traceback kernels)"},
        "dp_value": out["value"], "recombinations": [int(len(out["p1_edges"])) - 1, int(len(out["p2_edges"])) - 1],
        "gpu_launches": int(st["launches"]) * args.steps,
        "kernel_ms": {"sweep": sweep_ms, "traceback": float(np.mean(trace))},
        "roofline": {"bound": "hbm", "kernel": "dip_sweep4_kernel" if st["engine"] == 4 else "dip_sweep_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                     "cells_written_frac": float(st["cells_written"]) / max(1.0, float(st["cells"])),
                     "traffic_source": "no ncu capture for this workload"},
        "e2e": {"value": U / e2e_s, "unit": "cell-updates/s", "ms_per_step": e2e_s * 1e3, "samples_per_step": 1,
                "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(ctypes_out_bytes()),
                "api": "dg_dp_diploid (host buffers -> planning -> H2D -> program build -> sweep -> traceback -> D2H)"},
        "clocks": clk.summary(),
    }
    if not args.no_cpu_baseline and os.path.exists(REF_PLAIN):
        try:
            with tempfile.TemporaryDirectory() as td:
                gp = os.path.join(td, "graph.dgd")
                graph_to_dgd(g, gp)
                threads = min(os.cpu_count() or 8, 16)
                first = run_ref_dp(gp, args.R, threads, max_levels=min(g.n_levels, 300))
                per_level = first["ms"] / first["levels"]
                max_levels = 0 if per_level * g.n_levels <= 25000 else max(300, int(25000 / per_level))
                r = run_ref_dp(gp, args.R, threads, max_levels=max_levels)
                line["cpu_baseline"] = {"value": r["cell_updates"] / (r["ms"] * 1e-3), "unit": "cell-updates/s", "cores": threads, "kind": "reference",
                                        "sample": ("all %d levels" % g.n_levels) if not max_levels else ("first %d of %d levels" % (max_levels, g.n_levels)),
                                        "ms": r["ms"], "host_cpus": os.cpu_count()}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))
    p.close()
    ctx.close()
    return 0


def haploid_arm(args):
    """BASELINE config 1's DP: haploid DP + R+1 tracebacks on the Kahn-ordered expanded graph of MHC_4 +
    CHM13 reads (fixture from the reference front end).  U = (R+1) * nE cell-updates, B = (R+1) * (8 n + 8 nE)."""
    import torch
    import oracle
    from dipgenie_b200.cuda_api import Context, HapGraph
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the device path has no CPU fallback")
    torch.cuda.set_device(0)
    g = HapGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_hapin.npz"))
    R = args.R
    ctx = Context(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e2e_t = []
    for i in range(max(1, args.e2e_steps) + 1):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        o1 = ctx.dp_haploid(g, R)
        if i > 0:
            e2e_t.append(time.perf_counter() - t0)
    p = ctx.hap_create(g, R)
    ms, sweep, trace = [], [], []
    with ClockSampler(0) as clk:
        for i in range(args.warmup + args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            p.run()
            out = p.result()
            st = p.stats()
            if i >= args.warmup:
                ms.append(st["sweep_ms"] + st["traceback_ms"]); sweep.append(st["sweep_ms"]); trace.append(st["traceback_ms"])
    assert np.array_equal(out["colours_by_r"], o1["colours_by_r"])
    peak, peak_src = peaks()
    nE, n = len(g.adj_dst), g.n
    U, algo = float((R + 1) * nE), float((R + 1) * (8 * n + 8 * nE))
    dev_ms, sweep_ms, e2e_s = float(np.mean(ms)), float(np.mean(sweep)), float(np.mean(e2e_t))
    in_bytes = sum(int(np.asarray(a).nbytes) for a in (g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val))
    line = {
        "metric": "dp_cell_updates_per_sec", "value": U / (dev_ms * 1e-3), "unit": "cell-updates/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "bundled MHC_4 test panel + CHM13 reads (Kahn-ordered expanded graph fixture)",
        "config": {"workload": args.workload, "description": "haploid -p1 DP + R+1 tracebacks (approximator.cpp:44-168)", "R": R, "ploidy": 1,
                   "vertices": n, "edges": nE, "levels": st["n_levels"], "max_width": st["max_width"], "samples_per_step": 1,
                   "l2": "256 MiB device buffer rewritten between timed iterations", "timing": "CUDA events on the library stream (sweep + traceback kernels)"},
        "colours_by_r": [int(x) for x in out["colours_by_r"][:3]], "gpu_launches": int(st["launches"]) * args.steps,
        "kernel_ms": {"sweep": sweep_ms, "traceback": float(np.mean(trace))},
        "roofline": {"bound": "hbm", "kernel": "hap_sweep_kernel", "achieved": algo / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / (sweep_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                     "note": "one persistent CTA walks 120 363 dependent levels of 1-13 vertices: bound by the per-level barrier chain, not by HBM (DESIGN.md 5)"},
        "e2e": {"value": U / e2e_s, "unit": "cell-updates/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": int(sum(len(x) for x in o1["paths"]) * 4 + (R + 1) * 4), "api": "dg_dp_haploid (host buffers)"},
        "clocks": clk.summary(),
    }
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        ref = oracle.dp_haploid(g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.n_colours, R)
        dt = time.perf_counter() - t0
        assert np.array_equal(np.asarray(ref["colours_by_r"]), out["colours_by_r"])
        line["cpu_baseline"] = {"value": U / dt, "unit": "cell-updates/s", "cores": 1, "kind": "port", "sample": "the whole sample (oracle/dp_haploid.c, serial like the reference)", "ms": dt * 1e3}
    print(json.dumps(line))
    p.close(); ctx.close()
    return 0


def sketch_arm(args):
    """canonical (w,k)-minimizer sketch of every read (spectrum, per-hash read counts) and of every
    walk (joined with the spectrum, covered vertices) of MHC_4 + CHM13 reads.  Metric: input bases per second;
    B = 1 B per base + 8 B per emitted hash + 16 B per walk minimizer probed + 4 B per covered vertex id."""
    import torch
    import oracle
    from dipgenie_b200.cuda_api import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the device path has no CPU fallback")
    torch.cuda.set_device(0)
    z = dict(np.load(os.path.join(GOLD, "sketch_mhc4_chm13.npz")))
    k, w = int(z["k"]), int(z["w"])
    ctx = Context(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    walk_bases = int(sum(int(z["seg_off"][v + 1]) - int(z["seg_off"][v]) for v in z["walk_vtx"].tolist()))
    bases = int(len(z["read_bases"])) + walk_bases
    wall, kern, launches = [], [], 0
    with ClockSampler(0) as clk:
        for i in range(args.warmup + args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            t0 = time.perf_counter()
            sp, rcnt = ctx.sketch_reads(z["read_bases"], z["read_off"], k, w)
            s1 = ctx.sketch_stats()
            ix = ctx.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, sp)
            s2 = ctx.sketch_stats()
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                wall.append(dt); kern.append(s1["kernel_ms"] + s2["kernel_ms"]); launches += s1["launches"] + s2["launches"]
    n_min = int(s1["minimizers"]) + int(np.sum(ix["n_minimizers"]))
    algo = float(bases + 8 * n_min + 16 * int(np.sum(ix["n_minimizers"])) + 4 * len(ix["hit_vtx"]))
    peak, peak_src = peaks()
    kms, e2e_s = float(np.mean(kern)), float(np.mean(wall))
    h2d = int(z["read_bases"].nbytes + z["read_off"].nbytes + z["seg_bases"].nbytes + z["seg_off"].nbytes + z["walk_vtx"].nbytes + z["walk_off"].nbytes + z["top_order_map"].nbytes + sp.nbytes)
    d2h = int(sp.nbytes + rcnt.nbytes + sum(int(ix[x].nbytes) for x in ("hit_off", "hit_sid", "hit_vtx_off", "hit_vtx")))
    line = {
        "metric": "sketch_bases_per_sec", "value": bases / (kms * 1e-3), "unit": "bases/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "bundled MHC_4 test panel (5 walks) + CHM13 reads",
        "config": {"workload": args.workload, "description": "dg_sketch_reads + dg_index_walks (solver.cpp:277-446,526-576)", "k": k, "w": w,
                   "bases": bases, "read_bases": int(len(z["read_bases"])), "walk_bases": walk_bases, "minimizers": n_min, "spectrum": int(len(sp)),
                   "hits": int(len(ix["hit_sid"])), "l2": "256 MiB device buffer rewritten between timed iterations",
                   "timing": "value: CUDA events around the kernels of the two calls (the calls take host buffers: inputs are uploaded first, the timed kernels read them from HBM); e2e: wall clock of the two calls"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "sketch_tile_kernel + cover / spectrum / table kernels", "achieved": algo / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / (kms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo},
        "e2e": {"value": bases / e2e_s, "unit": "bases/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "dg_sketch_reads + dg_index_walks (host buffers)"},
        "clocks": clk.summary(),
    }
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        osp, _ = oracle.sketch_reads(z["read_bases"], z["read_off"], k, w)[:2]
        one = dict(walk_vtx=z["walk_vtx"][int(z["walk_off"][0]):int(z["walk_off"][1])], walk_off=np.array([0, int(z["walk_off"][1]) - int(z["walk_off"][0])], np.uint64))
        oracle.index_walks(z["seg_bases"], z["seg_off"], one["walk_vtx"], one["walk_off"], z["top_order_map"], k, w, osp)
        dt = time.perf_counter() - t0
        assert np.array_equal(np.asarray(osp), sp)
        sample_bases = int(len(z["read_bases"])) + int(sum(int(z["seg_off"][v + 1]) - int(z["seg_off"][v]) for v in one["walk_vtx"].tolist()))
        line["cpu_baseline"] = {"value": sample_bases / dt, "unit": "bases/s", "cores": 1, "kind": "port",
                                "sample": "all reads + the first walk (oracle/sketch.c)", "ms": dt * 1e3}
    print(json.dumps(line))
    ctx.close()
    return 0


def row_sharded_arm(args):
    """ONE diploid DP over the N GPUs of the node (`--mode row-sharded`, under torchrun): wide transitions split by destination
    row over world x ctas CTAs, rows and barrier arrivals exchanged inside the sweep kernel over NVLink peer mappings
    (dipgenie_b200.shard.RowShardedDip; no NCCL on the data path).  Strong scaling: the line also carries the single-GPU time
    of the same problem, and the run asserts that every rank's result is identical to it."""
    import torch
    import torch.distributed as dist
    from dipgenie_b200.cuda_api import Context
    from dipgenie_b200.shard import RowShardedDip
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world < 2:
        raise SystemExit("bench.py --mode row-sharded needs torchrun with at least 2 ranks")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload if args.workload.startswith("lanes") else "lanes640x24"
    g, desc = load_workload(name)
    ctx = Context(local)
    single = ctx.dip_create(g, args.R)
    t1 = []
    for _ in range(args.warmup + args.steps):
        single.run(); ref = single.result(); st1 = single.stats(); t1.append(st1["sweep_ms"] + st1["traceback_ms"] + st1["delta_ms"])
    single.close()
    prob = RowShardedDip(ctx, g, args.R, dist)
    tn = []
    for _ in range(args.warmup + args.steps):
        out = prob.run(); st = prob.stats(); tn.append(st["sweep_ms"] + st["traceback_ms"] + st["delta_ms"])
    good = bool(out["value"] == ref["value"] and out["s_het"] == ref["s_het"] and np.array_equal(out["p1_edges"], ref["p1_edges"])
                and np.array_equal(out["p2_edges"], ref["p2_edges"]))
    t = torch.tensor([float(np.mean(tn[args.warmup:])), float(np.mean(t1[args.warmup:])), 0.0 if good else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_sharded, ms_single, bad = [float(x) for x in t.tolist()]
    prob.close()
    if rank == 0:
        U = float(st["cell_updates"])
        peak, peak_src = peaks()
        print(json.dumps({
            "metric": "dp_cell_updates_per_sec", "value": U / (ms_sharded * 1e-3), "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_sharded, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "mode": "row-sharded",
            "config": {"workload": name, "description": desc, "R": args.R, "levels": st["n_levels"], "max_width": st["max_width"],
                       "ctas_per_rank": st["grid_ctas"], "engine_sharded": st["engine"], "engine_single_gpu": st1["engine"],
                       "single_gpu_ms": ms_single, "identical_on_every_rank": bad == 0.0,
                       "exchange": "in-kernel stores through NVLink peer mappings + system-scope counter barrier; no NCCL on the data path"},
            "roofline": {"bound": "hbm", "kernel": "dip_sweep_kernel (row-sharded)", "achieved": float(st["algo_bytes"]) / (ms_sharded * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": float(st["algo_bytes"]) / (ms_sharded * 1e-3) / 1e9 / (peak * world),
                         "traffic": None, "peak_source": peak_src + " x n_gpus"}}))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    return 0 if bad == 0.0 else 1


def load_samples(name: str, n: int, seed: int = 2):
    """n samples of the workload's shape.  Sample 0 is the fixture itself; the others keep its panel graph (every sample of
    the study is genotyped against a panel of this shape) and get their own read-derived part: the colour sets are dealt to
    the coloured vertices in a seeded permutation and the hom/het flags are redrawn at the same ratio — the levels differ in
    which cells carry scores, so values, predecessor codes and paths differ from sample to sample."""
    from dipgenie_b200.cuda_api import LevelGraph
    g, desc = load_workload(name)
    out = [g]
    ncol = np.diff(g.col_off)
    coloured = np.nonzero(ncol > 0)[0]
    for i in range(1, n):
        rng = np.random.default_rng(seed * 1000 + i)
        perm = rng.permutation(len(coloured))
        new_n = np.zeros_like(ncol)
        new_n[coloured] = ncol[coloured[perm]]
        off = np.concatenate([[0], np.cumsum(new_n)])
        val = np.empty(int(off[-1]), np.int32)
        for dst, src in zip(coloured, coloured[perm]):
            val[off[dst]:off[dst + 1]] = g.col_val[g.col_off[src]:g.col_off[src + 1]]
        hom = rng.permutation(g.colour_is_hom)
        out.append(LevelGraph(g.level_off, g.adj_off, g.adj_dst, g.adj_w, off, val, hom))
    return out, desc


def graph_to_dgd(g, path):
    from dipgenie_b200 import dgd
    dgd.save(path, dict(level_off=g.level_off, adj_off=g.adj_off, adj_dst=g.adj_dst, adj_w=g.adj_w, col_off=g.col_off,
                        col_val=g.col_val, colour_is_hom=g.colour_is_hom))


def run_ref_dp(graph_path, R, threads, max_levels=0, timeout=1800):
    cmd = [REF_PLAIN, "-G", graph_path, "-R", str(R), "-t", str(threads)]
    if max_levels:
        cmd += ["-M", str(max_levels)]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    m = re.search(r"DPONLY levels (\d+) threads (\d+) R (\d+) cell_updates (\d+) ms ([\d.]+)", p.stdout)
    if not m:
        raise RuntimeError(f"reference DP run failed: {p.stdout[-500:]} {p.stderr[-500:]}")
    return dict(levels=int(m.group(1)), threads=int(m.group(2)), cell_updates=int(m.group(4)), ms=float(m.group(5)))


def pick_ref_threads(graph_path, R, n_levels):
    """The reference's level loop pays 5 OpenMP barriers per level, so more threads is not always faster:
    probe a 4000-level sample with a few thread counts and keep the fastest (stated in the JSON)."""
    ncpu = os.cpu_count() or 1
    cands = sorted({min(ncpu, t) for t in (8, 16, 32, ncpu)})
    best = None
    probe = min(4000, n_levels)
    for t in cands:
        try:
            r = run_ref_dp(graph_path, R, t, max_levels=probe, timeout=600)
        except Exception:
            continue
        if best is None or r["ms"] < best[1]:
            best = (t, r["ms"])
    return (best[0] if best else min(ncpu, 8)), probe


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 5 + i and r[5 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    g, desc = load_workload(args.workload)
    if not os.path.exists(REF_PLAIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver_plain not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    with tempfile.TemporaryDirectory() as td:
        gp = os.path.join(td, "graph.dgd")
        graph_to_dgd(g, gp)
        threads, probe = pick_ref_threads(gp, args.R, g.n_levels)
        # bound each step to roughly <= 12 s of CPU wall time
        first = run_ref_dp(gp, args.R, threads, max_levels=min(g.n_levels, 20000))
        per_level = first["ms"] / first["levels"]
        max_levels = 0 if per_level * g.n_levels <= 12000 else max(2000, int(12000 / per_level))
        for _ in range(max(args.warmup - 1, 0)):
            run_ref_dp(gp, args.R, threads, max_levels=max_levels)
        ms, U = [], None
        for _ in range(args.steps):
            r = run_ref_dp(gp, args.R, threads, max_levels=max_levels)
            ms.append(r["ms"])
            U = r["cell_updates"]
        t = float(np.mean(ms))
        val = U / (t * 1e-3)
        sample = ("all %d levels" % g.n_levels) if not max_levels else ("first %d of %d levels" % (max_levels, g.n_levels))
        line = {
            "impl": "reference", "metric": "dp_cell_updates_per_sec", "value": val, "unit": "cell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "bundled MHC_4 test panel (levelized graph fixture), synthetic only where named",
            "config": {"workload": args.workload, "description": desc, "R": args.R, "ploidy": 2, "levels": int(g.n_levels),
                       "vertices": int(g.level_off[-1]), "max_width": int(np.diff(g.level_off).max()),
                       "cell_updates_per_sample": float(U) if not max_levels else None,
                       "reference_fn": "Approximator::diploid_dp_approximation_solver (src/approximator.cpp:362), unmodified objects, -G mode of oracle/ref_driver"},
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": threads, "kind": "reference", "sample": sample,
                             "host_cpus": os.cpu_count(), "thread_probe": f"fastest of {{8,16,32,nproc}} on a {probe}-level sample"},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="mhc4_chm13")
    ap.add_argument("--R", type=int, default=18)
    ap.add_argument("--samples-per-gpu", type=int, default=256, help="samples resident together on one GPU in the timed step (about 0.53 GB of HBM each)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct samples among the resident ones and in the batch call (see load_samples)")
    ap.add_argument("--ctas-per-sample", type=int, default=1, help="sweep CTAs per resident sample")
    ap.add_argument("--batch-ctas", type=int, default=0, help="CTAs per sample of the end-to-end batch call (0 = library default)")
    ap.add_argument("--batch", type=int, default=22, help="samples per GPU in the end-to-end batch call (22-sample study)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--mode", default="samples", help="samples (default: independent samples per GPU) | row-sharded (one DP over all GPUs)")
    ap.add_argument("--skip-e2e", action="store_true", help="diagnostics: no host-buffer batch calls before the resident group (the line then carries no e2e)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    if args.impl == "reference":
        return reference_arm(args)
    if args.mode == "row-sharded":
        return row_sharded_arm(args)
    if args.workload.startswith("c4_"):
        return wide_panel_arm(args)
    if args.workload == "mhc4_chm13_hap":
        return haploid_arm(args)
    if args.workload == "sketch_mhc4":
        return sketch_arm(args)

    import torch
    import torch.distributed as dist

    from dipgenie_b200.cuda_api import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the device path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # one process per GPU: this rank's share of the host cores for the planning threads of the batch call
    os.environ.setdefault("DG_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))))
    samples, desc = load_samples(args.workload, max(1, args.distinct), seed=2 + rank)
    g = samples[0]
    ctx = Context(local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # end-to-end through the host-buffer C-ABI batch call (planning, H2D, kernels, D2H inside the timed region)
    graphs = [samples[i % len(samples)] for i in range(max(1, args.batch))]
    e2e_t, single_t = [], []
    E2E_WARMUP = 2      # untimed calls: the library pins its page-locked plan blocks and fills its device pool during the first ones
    for i in range(0 if args.skip_e2e else args.e2e_steps + E2E_WARMUP):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ctx.dp_diploid_batch(graphs, args.R, ctas_per_sample=args.batch_ctas)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        o1 = ctx.dp_diploid(g, args.R)
        d1 = time.perf_counter() - t0
        if i >= E2E_WARMUP:
            e2e_t.append(dt)
            single_t.append(d1)
        assert res[0]["value"] == o1["value"]
        if i == 0:
            batch_values = [r["value"] for r in res]
        assert [r["value"] for r in res] == batch_values
    if args.skip_e2e:
        e2e_t, single_t, batch_values = [float("nan")], [float("nan")], []
        o1 = ctx.dp_diploid(g, args.R)
    e2e_value = o1["value"]
    e2e_s, single_s = float(np.mean(e2e_t)), float(np.mean(single_t))
    print("bench: e2e batch calls (s): %s; single-sample calls (s): %s" % ([round(x, 3) for x in e2e_t], [round(x, 3) for x in single_t]), file=sys.stderr)


    ctx.release_cached_memory()      # the batch calls' device blocks go back: the resident group needs most of the HBM
    # device-resident throughput: S samples in HBM (about 1 GB each), swept together
    S = max(1, args.samples_per_gpu)
    probs = [ctx.dip_create(samples[i % len(samples)], args.R, slot=i % 1024, ctas=args.ctas_per_sample) for i in range(S)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """One pass of the hot path over one batch: S resident samples, each its own persistent sweep."""
        flush.fill_(1)                      # evict L2 between iterations (not timed: timing is CUDA events in the library)
        torch.cuda.synchronize()
        ms = ctx.dip_run_many(probs)        # fork/join events on the library stream
        outs = [p.result() for p in probs]
        sts = [p.stats() for p in probs]
        return ms, outs, sts

    for _ in range(args.warmup):
        ms, outs, sts = step()
    barrier()
    group_ms, sweep, trace, delta = [], [], [], []
    t_wall0 = time.perf_counter()
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            ms, outs, sts = step()
            group_ms.append(ms)
            sweep += [st["sweep_ms"] for st in sts]
            trace += [st["traceback_ms"] for st in sts]
            delta += [st["delta_ms"] for st in sts]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    st = sts[0]
    dev_ms = float(np.mean(group_ms))
    assert outs[0]["value"] == e2e_value
    assert all(o["value"] == batch_values[i % len(samples)] for i, o in enumerate(outs) if i % len(samples) < len(batch_values))

    tmax = torch.tensor([dev_ms, e2e_s * 1e3, single_s * 1e3, float(np.mean(sweep))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, single_ms_max, sweep_ms_max = [float(x) for x in tmax.tolist()]

    if rank == 0:
        B = len(graphs)
        peak, peak_src = peaks()
        sweep_ms = float(np.mean(sweep))
        # dg_dip_run_many sweeps all S resident samples in ONE launch (dip_sweep_many_kernel) when it can: that launch's
        # algorithmic bytes are S samples' worth
        fused = S >= 2 and sts[1]["launches"] < sts[0]["launches"]
        per_launch = S if fused else 1
        algo = float(np.mean([x["algo_bytes"] for x in sts]))
        U = float(np.mean([x["cell_updates"] for x in sts]))
        achieved = per_launch * algo / (sweep_ms * 1e-3) / 1e9
        traffic = NCU_TRAFFIC.get((args.workload, args.R, S))
        kname = ("dip_sweep4_many_kernel" if fused else "dip_sweep4_kernel") if st["engine"] == 4 else ("dip_sweep_many_kernel" if fused else "dip_sweep_kernel")
        line = {
            "metric": "dp_cell_updates_per_sec", "value": world * S * U / (dev_ms_max * 1e-3), "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "bundled MHC_4 test panel (levelized graph fixture), synthetic only where named",
            "config": {"workload": args.workload, "description": desc, "R": args.R, "ploidy": 2, "levels": st["n_levels"],
                       "vertices": st["n_vertices"], "max_width": st["max_width"], "cell_updates_per_sample": U,
                       "dest_cells_per_sample": st["cells"], "samples_per_step": world * S, "samples_per_gpu": S,
                       "ctas_per_sample": st["grid_ctas"], "distinct_samples": len(samples), "engine": st["engine"],
                       "sharding": "independent samples per GPU and per CTA, no collective", "fused_sweep_launch": bool(fused),
                       "l2": "256 MiB device buffer rewritten between timed iterations",
                       "timing": "CUDA events on the library stream around the fork/join of the S resident sweeps (delta + sweep + traceback kernels)"},
            "samples_per_sec": world * B / (e2e_ms_max * 1e-3),
            "dp_value": outs[0]["value"],
            "gpu_launches": int(sum(x["launches"] for x in sts)) * args.steps,
            "kernel_ms": {"pair_scores": float(np.mean(delta)), "sweep": sweep_ms, "traceback": float(np.mean(trace)),
                          "note": ("sweep: the one fused launch over all S samples; traceback: the four launches over all S samples"
                                   if fused else "per launch, with S launches resident together")},
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic[0] * per_launch if traffic else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo * per_launch, "samples_per_launch": per_launch,
                         "launches_resident_together": 1 if fused else S,
                         "aggregate_achieved": achieved * (1 if fused else S),
                         "traffic_source": traffic[1] if traffic else "no ncu capture for this (workload, R, samples per GPU)",
                         "cells_written_frac": float(np.mean([x["cells_written"] for x in sts])) / max(1.0, float(np.mean([x["cells"] for x in sts]))),
                         "note": "algorithmic bytes (SURVEY 8d) count an HBM round trip of every cell of every level; the in-place layers "
                                 "leave the passive pairs untouched and keep the narrow levels in shared memory, so achieved/peak above 1 is possible; "
                                 "H = 5 samples are chains of ~10^5 dependent level transitions and the GPU is filled with samples side by side (DESIGN.md)"},
            "e2e": {"value": world * B * U / (e2e_ms_max * 1e-3), "unit": "cell-updates/s", "ms_per_step": e2e_ms_max,
                    "samples_per_step": world * B, "h2d_bytes_per_step": int(np.mean([x["h2d_bytes"] for x in sts])) * B,
                    "d2h_bytes_per_step": int(ctypes_out_bytes()) * B,
                    "h2d_note": "the gather-form tables dg_dip_create uploads (in-edge CSR, colour masks, slot / class tables, headers, directories); the level programs are built on the device",
                    "api": "dg_dp_diploid_batch (host buffers -> planning -> H2D -> program build -> sweep -> traceback -> D2H), %d samples per GPU" % B},
            "e2e_single_sample": {"value": U / (single_ms_max * 1e-3), "unit": "cell-updates/s", "ms_per_step": single_ms_max,
                                  "api": "dg_dp_diploid, one sample, whole GPU"},
            "clocks": clk.summary(),
        }
        if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_PLAIN):
            try:
                with tempfile.TemporaryDirectory() as td:
                    gp = os.path.join(td, "graph.dgd")
                    graph_to_dgd(g, gp)
                    threads, probe = pick_ref_threads(gp, args.R, g.n_levels)
                    first = run_ref_dp(gp, args.R, threads, max_levels=min(g.n_levels, 20000))
                    per_level = first["ms"] / first["levels"]
                    max_levels = 0 if per_level * g.n_levels <= 25000 else max(2000, int(25000 / per_level))
                    r = run_ref_dp(gp, args.R, threads, max_levels=max_levels)
                    line["cpu_baseline"] = {
                        "value": r["cell_updates"] / (r["ms"] * 1e-3), "unit": "cell-updates/s", "cores": threads, "kind": "reference",
                        "sample": ("one sample, all %d levels" % g.n_levels) if not max_levels else ("one sample, first %d of %d levels" % (max_levels, g.n_levels)),
                        "ms": r["ms"], "host_cpus": os.cpu_count()}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    for p in probs:
        p.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def ctypes_out_bytes():
    import ctypes
    from dipgenie_b200.cuda_api import DipOutput
    return ctypes.sizeof(DipOutput)


if __name__ == "__main__":
    sys.exit(main())
