"""GPU parity tests of the diploid DP (through the C ABI of libdipgenie_cuda.so, via ctypes):
CUDA path vs the oracle on the same seeded inputs (bit-exact: value, s_het, edge lists and the
per-level checksums of every DP layer), vs the reference's committed goldens, and — at full size —
vs the reference's digest of all 120 362 layers of the bundled MHC data."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD, assert_dip_equal, oracle_dip
from dipgenie_b200 import dgd, synth
from dipgenie_b200.cuda_api import Context, LevelGraph

pytestmark = pytest.mark.gpu

TINY_DIP = ["test_p2_R2_k5_w3", "test_p2_R0_k3_w2", "test_p2_R1_k3_w2", "test_p2_R2_k3_w2", "test2_p2_R2"]


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def cuda_dip(ctx, g, R, checks=True):
    p = ctx.dip_create(g, R)
    try:
        p.run(checksums=checks)
        out = p.result()
        if checks:
            out["checksum"], out["live"] = p.checksums()
        out["stats"] = p.stats()
        return out
    finally:
        p.close()


@pytest.mark.parametrize("name", TINY_DIP)
def test_cuda_matches_reference_tiny(name, ctx, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = LevelGraph.from_dgd(d)
    R = int(d["dip_in.R"][0])
    o = cuda_dip(ctx, g, R)
    e = expected["tiny"][name]
    assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert np.array_equal(o["checksum"][1:], d["dip_out.level_checksum"][1:])
    one_shot = ctx.dp_diploid(g, R)
    assert one_shot["value"] == e["value"] and one_shot["p1_edges"].ravel().tolist() == e["p1_edges"]


@pytest.fixture(params=["v4", "v3"])
def engine(request, monkeypatch):
    """The level-program engine (default) or, with DG_ENGINE_V3=1, the task-stream engine that takes the problems
    outside the packed-key program's range."""
    if request.param == "v3":
        monkeypatch.setenv("DG_ENGINE_V3", "1")
    return request.param


@pytest.mark.parametrize("seed", range(24))
def test_cuda_matches_oracle_random(seed, ctx, oracle_mod, engine):
    rng = np.random.default_rng(2000 + seed)
    g = synth.random_level_graph(seed, n_levels=int(rng.integers(2, 40)), max_width=int(rng.integers(1, 24)),
                                 n_colours=int(rng.integers(0, 200)), p_weight1=float(rng.random() * 0.6),
                                 p_colour=float(rng.random()), max_out=int(rng.integers(1, 5)))
    R = int(rng.integers(0, 9))
    o = cuda_dip(ctx, g, R)
    assert_dip_equal(oracle_dip(oracle_mod, g, R), o)
    assert o["stats"]["engine"] in ((3,) if engine == "v3" else (3, 4))     # (a cell of more than 1024 candidates: not the program's)


@pytest.mark.parametrize("variant", [dict(DG_V4_SLOG="9"), dict(DG_V4_RC="5"), dict(DG_V4_SLOT="256"), dict(DG_V4_NCW="3"),
                                     dict(DG_V4_SLOG="9", DG_V4_RC="5", DG_V4_NCW="16")])
def test_cuda_level_program_kernel_variants(variant, ctx, oracle_mod, monkeypatch):
    """Narrower shared-memory layers (stride 512: levels wider than 22 go to HBM), 5 layers per register chunk, ring slots
    too small for a program (descriptors read in place from HBM), few / many compute warps."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    n4 = 0
    for seed in range(10):
        rng = np.random.default_rng(5000 + seed)
        if seed % 4 == 3:
            g = synth.lane_panel_graph(seed, n_lanes=int(rng.integers(8, 31)), n_blocks=4, rec_per_block=2, p_colour=0.25, n_colours=300)
        else:
            g = synth.random_level_graph(700 + seed, n_levels=int(rng.integers(2, 50)), max_width=int(rng.integers(1, 40)),
                                         n_colours=int(rng.integers(0, 200)), p_weight1=float(rng.random() * 0.6),
                                         p_colour=float(rng.random()), max_out=int(rng.integers(1, 5)))
        R = int(rng.integers(0, 12))
        o = cuda_dip(ctx, g, R)
        n4 += o["stats"]["engine"] == 4
        assert_dip_equal(oracle_dip(oracle_mod, g, R), o)
    assert n4 >= 8


def test_device_built_program_equals_host_built(ctx, dp_emu4):
    """prog_fill_kernel against prog_fill_level_host (the same dp_prog.h functions, run serially): byte for byte."""
    cases = [(synth.random_level_graph(300 + s, n_levels=30, max_width=36, n_colours=150, p_colour=0.5, max_out=4), 4) for s in range(4)]
    cases.append((synth.lane_panel_graph(5, n_lanes=24, n_blocks=5, rec_per_block=3, p_colour=0.3, n_colours=2000), 7))
    cases.append((LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))[0], 18))
    n4 = 0
    for g, R in cases:
        p = ctx.dip_create(g, R)
        try:
            dev = p.debug_program()
            st = p.stats()
        finally:
            p.close()
        if st["engine"] != 4:           # (a cell of more than 1024 candidates: not the program's)
            continue
        n4 += 1
        host = dp_emu4.build_program(g, R, shape=(10, 31, 8192, 1, 10))
        assert st["prog_bytes"] == len(host) == len(dev)
        assert np.array_equal(dev, host)


def test_cuda_edge_cases(ctx, oracle_mod):
    one = LevelGraph([0, 1], [0, 0], [], [], [0, 0], [], [0])
    assert_dip_equal(oracle_dip(oracle_mod, one, 3), cuda_dip(ctx, one, 3))
    dead = LevelGraph([0, 1, 2, 3], [0, 1, 2, 2], [1, 2], [1, 1], [0, 0, 0, 0], [], [0])
    for R in (0, 3, 4):
        assert_dip_equal(oracle_dip(oracle_mod, dead, R), cuda_dip(ctx, dead, R))
    k = 300   # fan-in above 255 -> 32-bit predecessor codes
    level_off = [0, 1, 1 + k, 2 + k, 3 + k]
    adj_off = [0, k] + list(range(k + 1, 2 * k + 1)) + [2 * k + 1, 2 * k + 1]
    adj_dst = list(range(1, 1 + k)) + [1 + k] * k + [2 + k]
    adj_w = [0] * k + [i % 2 for i in range(k)] + [0]
    ncol = np.zeros(3 + k, np.int64)
    ncol[1:1 + k] = 1
    g = LevelGraph(level_off, adj_off, adj_dst, adj_w, np.concatenate([[0], np.cumsum(ncol)]), np.arange(k) % 7,
                   [1, 0, 1, 0, 0, 1, 0])
    o = cuda_dip(ctx, g, 2)
    assert o["stats"]["pred_bytes"] == 4
    assert_dip_equal(oracle_dip(oracle_mod, g, 2), o)


def test_cuda_rejects_malformed_input(ctx):
    """dg_dip_create returns DG_ERR_ARG (no launch) for weights above 1 and for parallel edges of differing weight."""
    from dipgenie_b200.cuda_api import DipGenieCudaError
    heavy = LevelGraph([0, 1, 3, 4], [0, 2, 3, 4, 4], [1, 2, 3, 3], [0, 2, 0, 0], [0, 0, 0, 0, 0], [], [0])
    mixed = LevelGraph([0, 1, 2, 3], [0, 2, 3, 3], [1, 1, 2], [0, 1, 0], [0, 0, 0, 0], [], [0])
    for bad, what in ((heavy, "weight above 1"), (mixed, "parallel edges")):
        with pytest.raises(DipGenieCudaError, match=what):
            ctx.dip_create(bad, 2)
        with pytest.raises(DipGenieCudaError, match=what):
            ctx.dp_diploid(bad, 2)


def test_cuda_wide_levels_use_many_ctas(ctx, oracle_mod):
    """Widths large enough that transitions are spread over the whole grid and closed by the counter barrier."""
    g = synth.lane_panel_graph(11, n_lanes=72, n_blocks=8, rec_per_block=3, p_colour=0.2, n_colours=500)
    o = cuda_dip(ctx, g, 5)
    assert o["stats"]["grid_ctas"] > 1
    assert_dip_equal(oracle_dip(oracle_mod, g, 5), o)


def test_cuda_level_program_cells_beyond_the_packed_ordinal(ctx, oracle_mod):
    """Panels of 33 .. 181 walks: a recombination x recombination cell has in-degree^2 > 1024 candidates; the warp form keeps
    the round in the key and takes the lane from a ballot (dp_sweep4.cuh: big_cell), the code is the plain ordinal.  Value,
    paths and the checksums of every live cell of every level against the oracle, shared-memory and HBM-resident levels."""
    for seed, lanes, blocks, R in ((41, 40, 6, 4), (42, 90, 4, 7), (43, 128, 3, 2)):
        g = synth.lane_panel_graph(seed, n_lanes=lanes, n_blocks=blocks, rec_per_block=3, p_colour=0.25, n_colours=300)
        o = cuda_dip(ctx, g, R)
        assert o["stats"]["engine"] == 4
        assert_dip_equal(oracle_dip(oracle_mod, g, R), o)


@pytest.mark.parametrize("R", [18, 0, 6, 36])
def test_cuda_matches_reference_mhc_full_size(R, ctx, expected, engine):
    """BASELINE config 2 graph (MHC_4.gfa.gz, CHM13 reads): every DP layer digest-equal to the reference."""
    if engine == "v3" and R not in (18, 36):
        pytest.skip("task-stream engine: R = 18 and 36 only (time)")
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = cuda_dip(ctx, g, R)
    assert o["stats"]["engine"] == (3 if engine == "v3" else 4)
    e = expected["mhc4_chm13"]["diploid"][str(R)]
    assert o["value"] == e["value"]
    assert o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert hashlib.sha256(o["checksum"][1:].tobytes()).hexdigest() == e["checksum_sha256"]
    assert hashlib.sha256(o["live"][1:].tobytes()).hexdigest() == e["live_sha256"]
    # the unchecked (timed) kernel variant must give the same answer
    p = ctx.dip_create(g, R)
    p.run(checksums=False)
    o2 = p.result()
    p.close()
    assert o2["value"] == e["value"] and o2["p1_edges"].ravel().tolist() == e["p1_edges"]


def test_cuda_batch_of_samples_matches_one_by_one(ctx, oracle_mod, engine):
    """dg_dp_diploid_batch: samples resident together on their own streams / CTA groups give exactly the
    results of separate calls (shapes chosen so that narrow, wide and hand-over tasks all occur)."""
    graphs, Rs = [], []
    for seed in range(14):
        rng = np.random.default_rng(4000 + seed)
        if seed % 5 == 4:
            g = synth.lane_panel_graph(seed, n_lanes=int(rng.integers(8, 40)), n_blocks=4, rec_per_block=2, p_colour=0.25, n_colours=300)
        else:
            g = synth.random_level_graph(900 + seed, n_levels=int(rng.integers(2, 60)), max_width=int(rng.integers(1, 30)),
                                         n_colours=int(rng.integers(0, 120)), p_colour=float(rng.random()))
        graphs.append(g)
        Rs.append(int(rng.integers(0, 8)))
    for conc, ctas in ((0, 0), (3, 2), (1, 1)):
        outs = ctx.dp_diploid_batch(graphs, Rs, max_concurrent=conc, ctas_per_sample=ctas)
        for g, R, o in zip(graphs, Rs, outs):
            assert_dip_equal(oracle_dip(oracle_mod, g, R, want_checksums=False), o, checks=False)


def test_cuda_batch_mhc_samples(ctx, expected):
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    Rs = [18, 0, 6, 36, 18, 6]
    outs = ctx.dp_diploid_batch([g] * len(Rs), Rs)
    for R, o in zip(Rs, outs):
        e = expected["mhc4_chm13"]["diploid"][str(R)]
        assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
        assert o["p1_edges"].ravel().tolist() == e["p1_edges"] and o["p2_edges"].ravel().tolist() == e["p2_edges"]


def test_cuda_unpacked_arithmetic_path(ctx, oracle_mod, expected, monkeypatch):
    """DG_NO_PACK=1 forces the unpacked lane/pair arithmetic that problems with DP values >= 2^21 take
    (the default packs value and ordinals into one word per layer, dp_cell.h)."""
    monkeypatch.setenv("DG_NO_PACK", "1")
    for seed in range(8):
        rng = np.random.default_rng(7000 + seed)
        g = synth.random_level_graph(60 + seed, n_levels=int(rng.integers(2, 40)), max_width=int(rng.integers(1, 24)),
                                     n_colours=int(rng.integers(0, 200)), p_weight1=float(rng.random() * 0.6),
                                     p_colour=float(rng.random()), max_out=int(rng.integers(1, 5)))
        R = int(rng.integers(0, 9))
        assert_dip_equal(oracle_dip(oracle_mod, g, R), cuda_dip(ctx, g, R))
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = cuda_dip(ctx, g, 18)
    e = expected["mhc4_chm13"]["diploid"]["18"]
    assert o["value"] == e["value"] and o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert hashlib.sha256(o["checksum"][1:].tobytes()).hexdigest() == e["checksum_sha256"]


def test_cuda_destinations_with_more_than_32_in_edges(ctx, oracle_mod, monkeypatch):
    """Slice blocks + scratch combine (TK_LONG) for recombination vertices of panels with more than 32 walks, in
    shared-memory (small R) and HBM layers, with staged and in-place pair scores; DG_NO_LONG=1 forces the pair form."""
    from test_dp_diploid_cpu import funnel_graph
    for seed, lanes in ((21, 40), (22, 70), (23, 33), (24, 97)):
        g = funnel_graph(seed, lanes)
        for R in (0, 3, 12):
            ref = oracle_dip(oracle_mod, g, R)
            assert_dip_equal(ref, cuda_dip(ctx, g, R))
    g = funnel_graph(31, 64, n_funnels=3)
    ref = oracle_dip(oracle_mod, g, 5)
    monkeypatch.setenv("DG_NO_LONG", "1")
    assert_dip_equal(ref, cuda_dip(ctx, g, 5))


def test_records_larger_than_a_slot_run_in_lane_form(ctx, oracle_mod):
    """Panels of several hundred lanes: a level's record (offsets, in-edges, destinations, blocks) no longer fits a 4 KB
    task slot; the lane form then reads it in place from global memory instead of falling back to the pair form."""
    g = synth.lane_panel_graph(21, n_lanes=400, n_blocks=3, rec_per_block=3, p_colour=0.1, n_colours=4096)
    for R in (2, 9):
        o = cuda_dip(ctx, g, R)                      # value, paths and the per-level checksums of every layer
        assert_dip_equal(oracle_dip(oracle_mod, g, R), o)
        assert o["stats"]["n_wide"] > 0


def test_cuda_many_resident_problems_one_fused_launch(ctx, oracle_mod, monkeypatch, engine):
    """dg_dip_run_many starts problems that live in distinct slots together; more than the 32 hardware work queues'
    worth of them run as ONE sweep launch (dip_sweep_many_kernel: CTA -> (problem, local CTA) map).  Same results as
    one-by-one runs and as the per-stream launches (DG_NO_FUSED_MANY=1)."""
    graphs, Rs = [], []
    for seed in range(40):
        rng = np.random.default_rng(8100 + seed)
        if seed % 7 == 3:
            g = synth.lane_panel_graph(seed, n_lanes=int(rng.integers(8, 50)), n_blocks=3, rec_per_block=2, p_colour=0.25, n_colours=300)
        else:
            g = synth.random_level_graph(400 + seed, n_levels=int(rng.integers(2, 50)), max_width=int(rng.integers(1, 26)),
                                         n_colours=int(rng.integers(0, 120)), p_colour=float(rng.random()))
        graphs.append(g)
        Rs.append(int(rng.integers(0, 7)))
    want = [oracle_dip(oracle_mod, g, R, want_checksums=False) for g, R in zip(graphs, Rs)]
    probs = [ctx.dip_create(g, R, slot=i, ctas=1 + i % 3) for i, (g, R) in enumerate(zip(graphs, Rs))]
    try:
        for mode in ("fused", "streams", "fused"):
            if mode == "streams":
                monkeypatch.setenv("DG_NO_FUSED_MANY", "1")
            else:
                monkeypatch.delenv("DG_NO_FUSED_MANY", raising=False)
            ms = ctx.dip_run_many(probs)
            assert ms > 0
            for w, p in zip(want, probs):
                assert_dip_equal(w, p.result(), checks=False)
    finally:
        for p in probs:
            p.close()
