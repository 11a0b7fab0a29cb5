"""Row-sharded diploid DP (include/dipgenie_cuda.h: dg_dip_create_sharded; north_star "h1-row tiles"): the sweep of one
problem split by destination row over several ranks, rows and barrier arrivals exchanged inside the kernel.

* one GPU: all ranks as sibling problems of one process (dg_dip_attach_in_process) — the kernel logic (pushes,
  system-scope barriers, exit barrier) without IPC;
* two or more GPUs (skipped on a 1-GPU box): one process per GPU under torch.distributed, CUDA IPC peer mappings
  over NVLink (dipgenie_b200.shard.RowShardedDip)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from dipgenie_b200 import synth
from dipgenie_b200.cuda_api import Context, LevelGraph
from conftest import assert_dip_equal, oracle_dip
from test_dp_diploid_cpu import funnel_graph

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def run_siblings(ctx, g, R, world, ctas):
    probs = ctx.dip_sharded_in_process(g, R, world, ctas)
    try:
        outs = Context.run_sharded_siblings(probs)
        again = Context.run_sharded_siblings(probs)       # re-armed problems give the same answer
    finally:
        for p in probs:
            p.close()
    for a, b in zip(outs, again):
        assert_dip_equal(a, b, checks=False)
    return outs


@pytest.mark.parametrize("world,ctas", [(2, 4), (3, 2), (4, 8)])
def test_siblings_wide_panels_match_oracle(world, ctas, ctx, oracle_mod):
    # lane panels wider than the shared-memory tiles: nearly every transition is wide and row-split over the ranks
    for seed, lanes in ((3, 40), (4, 64)):
        g = synth.lane_panel_graph(seed, n_lanes=lanes, n_blocks=6, rec_per_block=2, p_colour=0.3, n_colours=256)
        want = oracle_dip(oracle_mod, g, 4, want_checksums=False)
        for o in run_siblings(ctx, g, 4, world, ctas):
            assert_dip_equal(want, o, checks=False)
    g = funnel_graph(9, lanes=48)                          # destinations with more than 32 in-edges
    want = oracle_dip(oracle_mod, g, 3, want_checksums=False)
    for o in run_siblings(ctx, g, 3, world, ctas):
        assert_dip_equal(want, o, checks=False)


def test_siblings_mixed_narrow_wide_random(ctx, oracle_mod):
    for seed in range(6):
        rng = np.random.default_rng(40 + seed)
        g = synth.random_level_graph(seed + 300, n_levels=int(rng.integers(20, 60)), max_width=int(rng.integers(20, 60)),
                                     n_colours=200, p_colour=0.4)
        R = int(rng.integers(1, 8))
        want = oracle_dip(oracle_mod, g, R, want_checksums=False)
        for o in run_siblings(ctx, g, R, 2, 6):
            assert_dip_equal(want, o, checks=False)


def test_siblings_mhc_golden(ctx):
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    exp = json.load(open(os.path.join(GOLD, "e2e_expected.json")))
    outs = run_siblings(ctx, g, 18, 2, 24)
    single = ctx.dp_diploid(g, 18)
    for o in outs:
        assert o["value"] == 60729
        assert_dip_equal(single, o, checks=False)
    assert exp is not None


def test_two_processes_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29621", os.path.join(ROOT, "tests", "perf", "run_row_sharded.py"), "--check"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "row-sharded OK" in p.stdout


def test_dead_peer_costs_one_timeout_not_one_per_level(ctx, monkeypatch):
    """A rank whose peer never runs gives up at its first cross-GPU wait, flags the run (sticky: no later wait spins,
    no further pushes) and dg_dip_result fails loudly — hundreds of wide levels must not cost hundreds of timeouts."""
    import time
    from dipgenie_b200.cuda_api import DipGenieCudaError
    monkeypatch.setenv("DG_SHARD_TIMEOUT_MS", "300")
    g = synth.lane_panel_graph(5, n_lanes=48, n_blocks=60, rec_per_block=2, p_colour=0.2, n_colours=256)   # ~240 wide levels
    probs = ctx.dip_sharded_in_process(g, 3, 2, 4)
    try:
        for p in probs:
            p.shard_arm()
        t0 = time.perf_counter()
        probs[0].run()                        # rank 1 is never launched
        with pytest.raises(DipGenieCudaError, match="timed out"):
            probs[0].result()
        assert time.perf_counter() - t0 < 5.0
    finally:
        for p in probs:
            p.close()
