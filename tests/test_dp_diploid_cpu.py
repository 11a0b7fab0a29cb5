"""CPU-side checks of the diploid DP path (no GPU needed):
  * the oracle (oracle/dp_diploid.c) against the reference's own results in tests/golden/
    (sink value, s_het, recombination-edge lists, sha256 of all per-level DP checksums);
  * the kernel-logic emulation (tests/emu: dp_prep.cpp + dp_cell.h, i.e. the code the CUDA kernels run)
    against the oracle on random levelized graphs, including the edge cases the domain has
    (R=0, dead sink, single transition, duplicate edges, fan-in > 255 -> 32-bit predecessor codes).
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD, assert_dip_equal, oracle_dip
from dipgenie_b200 import dgd, synth
from dipgenie_b200.cuda_api import LevelGraph

TINY_DIP = ["test_p2_R2_k5_w3", "test_p2_R0_k3_w2", "test_p2_R1_k3_w2", "test_p2_R2_k3_w2", "test2_p2_R2"]


@pytest.mark.parametrize("name", TINY_DIP)
def test_oracle_matches_reference_tiny(name, oracle_mod, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = LevelGraph.from_dgd(d)
    R = int(d["dip_in.R"][0])
    o = oracle_dip(oracle_mod, g, R)
    e = expected["tiny"][name]
    assert o["value"] == e["value"] == int(d["dip_out.value"][0])
    assert o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert np.array_equal(o["checksum"][1:], d["dip_out.level_checksum"][1:])
    assert np.array_equal(o["live"][1:], d["dip_out.level_live"][1:])


@pytest.mark.parametrize("name", TINY_DIP)
def test_emulated_kernel_matches_reference_tiny(name, dp_emu, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = LevelGraph.from_dgd(d)
    R = int(d["dip_in.R"][0])
    for f32 in (False, True):
        o = dp_emu.dp_diploid(g, R, force_pred32=f32)
        e = expected["tiny"][name]
        assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
        assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
        assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
        assert np.array_equal(o["checksum"][1:], d["dip_out.level_checksum"][1:])


@pytest.mark.parametrize("R", [18, 0])
def test_oracle_matches_reference_mhc(R, oracle_mod, expected):
    """Full-size pin: MHC_4.gfa.gz + CHM13 reads, all 120 362 DP layers (about 15 s per R)."""
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = oracle_dip(oracle_mod, g, R)
    e = expected["mhc4_chm13"]["diploid"][str(R)]
    assert o["value"] == e["value"]
    assert o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert hashlib.sha256(o["checksum"][1:].tobytes()).hexdigest() == e["checksum_sha256"]
    assert hashlib.sha256(o["live"][1:].tobytes()).hexdigest() == e["live_sha256"]


def test_emulated_kernel_matches_reference_mhc(dp_emu, expected):
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = dp_emu.dp_diploid(g, 6)
    e = expected["mhc4_chm13"]["diploid"]["6"]
    assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert hashlib.sha256(o["checksum"][1:].tobytes()).hexdigest() == e["checksum_sha256"]


@pytest.mark.parametrize("seed", range(40))
def test_emulated_kernel_matches_oracle_random(seed, oracle_mod, dp_emu):
    rng = np.random.default_rng(1000 + seed)
    g = synth.random_level_graph(seed, n_levels=int(rng.integers(2, 20)), max_width=int(rng.integers(1, 9)),
                                 n_colours=int(rng.integers(0, 90)), p_weight1=float(rng.random() * 0.6),
                                 p_colour=float(rng.random()), max_out=int(rng.integers(1, 5)))
    R = int(rng.integers(0, 7))
    assert_dip_equal(oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R))


def test_single_level_and_single_transition(oracle_mod, dp_emu):
    one = LevelGraph([0, 1], [0, 0], [], [], [0, 0], [], [0])
    assert_dip_equal(oracle_dip(oracle_mod, one, 3), dp_emu.dp_diploid(one, 3))
    two = LevelGraph([0, 1, 2], [0, 1, 1], [1], [1], [0, 1, 2], [0, 0], [1])
    for R in (0, 1, 2):
        assert_dip_equal(oracle_dip(oracle_mod, two, R), dp_emu.dp_diploid(two, R))


def test_dead_sink_when_R_too_small(oracle_mod, dp_emu):
    # source -w1-> a -w1-> sink : needs 4 recombinations for the pair of paths
    g = LevelGraph([0, 1, 2, 3], [0, 1, 2, 2], [1, 2], [1, 1], [0, 0, 0, 0], [], [0])
    for R in (0, 3, 4):
        a, b = oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R)
        assert_dip_equal(a, b)
    assert oracle_dip(oracle_mod, g, 3)["value"] < 0 and oracle_dip(oracle_mod, g, 4)["value"] == 0


def test_fan_in_above_255_uses_32bit_codes(oracle_mod, dp_emu):
    k = 300
    level_off = [0, 1, 1 + k, 2 + k, 3 + k]
    adj_off = [0, k] + list(range(k + 1, 2 * k + 1)) + [2 * k + 1, 2 * k + 1]
    adj_dst = list(range(1, 1 + k)) + [1 + k] * k + [2 + k]
    adj_w = [0] * k + [i % 2 for i in range(k)] + [0]
    ncol = np.zeros(3 + k, np.int64)
    ncol[1:1 + k] = 1
    col_off = np.concatenate([[0], np.cumsum(ncol)])
    col_val = np.arange(k) % 7
    g = LevelGraph(level_off, adj_off, adj_dst, adj_w, col_off, col_val, [1, 0, 1, 0, 0, 1, 0])
    for R in (0, 2):
        assert_dip_equal(oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R))


def test_lane_panel_model(oracle_mod, dp_emu):
    g = synth.lane_panel_graph(7, n_lanes=6, n_blocks=5, rec_per_block=2, p_colour=0.3, n_colours=64)
    for R in (0, 3):
        assert_dip_equal(oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R))


# (grid, threads, tile_cells, slot_bytes, delta_max_in, trace_T[, no_pack])
SHAPES = [
    (32, 480, 16384, 4096, 0, 0),       # the kernel's geometry
    (32, 480, 16384, 4096, 0, 0, 1),    # ... with the unpacked arithmetic (values of any size)
    (8, 64, 256, 512, 0, 5),            # small tiles and slots: narrow/wide hand-overs, row-split tasks
    (4, 32, 64, 160, 0, 3),             # slots too small for most records: in-place records
    (3, 16, 100000, 256, -1, 1),        # no matrices: on-the-fly masks everywhere (never narrow when coloured)
    (5, 32, 300, 200, 6, 2),            # matrices only for very few in-edges: mixed matrix / masks
]


@pytest.mark.parametrize("shape", SHAPES)
def test_all_sweep_modes_agree(shape, oracle_mod, dp_emu):
    """Narrow (shared-memory layers) and wide (row-split over CTAs, layers in HBM) transitions, staged and
    in-place records, staged / in-place / on-the-fly pair scores, and every hand-over between them, give the
    same layers, and the checkpointed traceback gives the same lists for any checkpoint distance."""
    seen = dict(narrow=0, wide=0, tasks=0, tasks_global=0, tasks_masks=0, matrices=0, tasks_lanes=0, tasks_long=0, tasks_lanes_inplace=0)
    for seed in range(12):
        rng = np.random.default_rng(77 + seed)
        g = synth.random_level_graph(500 + seed, n_levels=int(rng.integers(3, 30)), max_width=int(rng.integers(2, 12)),
                                     n_colours=int(rng.integers(0, 150)), p_colour=0.5)
        R = int(rng.integers(0, 6))
        o = dp_emu.dp_diploid(g, R, shape=shape)
        assert_dip_equal(oracle_dip(oracle_mod, g, R), o)
        for k in seen:
            seen[k] += o["modes"][k]
    if shape[0] in (32, 8):
        assert 0 < seen["tasks_lanes"] < seen["tasks"]    # both evaluation forms are exercised
    if shape[0] == 8:
        assert seen["narrow"] > 0 and seen["wide"] > 0 and seen["tasks"] > seen["narrow"] + seen["wide"]
    if shape[0] == 4:
        assert seen["tasks_global"] > 0
        assert seen["tasks_lanes_inplace"] > 0          # lane form reading a record that is too big for a slot in place
    if shape[0] == 3:
        assert seen["tasks_masks"] > 0 and seen["matrices"] == 0
    if shape[0] == 5:
        assert seen["tasks_masks"] > 0 and seen["matrices"] > 0


def test_mhc_mode_mix(dp_emu):
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = dp_emu.dp_diploid(g, 0)
    m = o["modes"]
    assert m["narrow"] + m["wide"] == g.n_levels - 1
    assert m["narrow"] > 0.9 * g.n_levels        # the bundled panel is almost entirely shared-memory resident
    assert m["tasks_global"] == 0 and m["tasks_masks"] == 0
    assert m["tasks_lanes"] > 0.95 * m["tasks"]           # nearly every task of real data takes the lane form


@pytest.mark.parametrize("R", [9, 12, 25])
def test_ten_layers_per_lane_variant(R, oracle_mod, dp_emu):
    """With packed keys and at least 10 layers the lane form keeps 10 layers per lane (LANE_RC_BIG)."""
    for seed in range(6):
        rng = np.random.default_rng(9100 + seed)
        g = synth.random_level_graph(800 + seed, n_levels=int(rng.integers(3, 25)), max_width=int(rng.integers(2, 14)),
                                     n_colours=int(rng.integers(0, 120)), p_weight1=float(rng.random() * 0.6), p_colour=float(rng.random()))
        assert_dip_equal(oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R))
    g = synth.lane_panel_graph(5, n_lanes=10, n_blocks=6, rec_per_block=2, p_colour=0.3, n_colours=96)
    assert_dip_equal(oracle_dip(oracle_mod, g, R), dp_emu.dp_diploid(g, R))


def funnel_graph(seed, lanes, n_funnels=2, n_colours=64):
    """source -> `lanes` lane vertices; per funnel: every lane also feeds (weight 1) one recombination vertex, which
    feeds every lane of the next level (weight 0) — a destination and a row with `lanes` in-edges; ... -> sink."""
    rng = np.random.default_rng(seed)
    level_sizes = [1, lanes]
    for _ in range(n_funnels):
        level_sizes += [lanes + 1, lanes]
    level_sizes += [1]
    level_off = np.concatenate([[0], np.cumsum(level_sizes)])
    adj = [[] for _ in range(level_off[-1])]
    for i in range(lanes):
        adj[0].append((level_off[1] + i, 0))
    l = 1
    for _ in range(n_funnels):
        a, b, c = level_off[l], level_off[l + 1], level_off[l + 2]
        x = b + lanes                                   # the recombination vertex, last position of its level
        for i in range(lanes):
            adj[a + i].append((b + i, 0))
            adj[a + i].append((x, 1))
            adj[b + i].append((c + i, 0))
            adj[x].append((c + i, 0))
        l += 2
    for i in range(lanes):
        adj[level_off[l] + i].append((level_off[l + 1], 0))
    adj_off = np.concatenate([[0], np.cumsum([len(x) for x in adj])])
    adj_dst = [v for x in adj for v, _ in x]
    adj_w = [w for x in adj for _, w in x]
    ncol = (rng.random(level_off[-1]) < 0.4).astype(np.int64) * rng.integers(1, 4, level_off[-1])
    ncol[0] = 0
    col_off = np.concatenate([[0], np.cumsum(ncol)])
    col_val = np.concatenate([np.sort(rng.choice(n_colours, int(c), replace=False)) for c in ncol] + [np.zeros(0, np.int64)])
    return LevelGraph(level_off, adj_off, adj_dst, adj_w, col_off, col_val, rng.integers(0, 2, n_colours))


@pytest.mark.parametrize("R", [0, 3, 12])
def test_destinations_with_more_than_32_in_edges(R, oracle_mod, dp_emu):
    """Panels with more than 32 walks: recombination vertices collect > 32 in-edges; the lane form cuts them into
    slice blocks and combines the slices through scratch words (TK_LONG); without it those levels take the pair form."""
    for seed, lanes in ((21, 40), (22, 70), (23, 33), (24, 97)):
        g = funnel_graph(seed, lanes)
        ref = oracle_dip(oracle_mod, g, R)
        o = dp_emu.dp_diploid(g, R)
        assert_dip_equal(ref, o)
        assert o["modes"]["tasks_long"] > 0
        o2 = dp_emu.dp_diploid(g, R, shape=(0, 0, 0, 0, 0, 0, 0, 1))
        assert_dip_equal(ref, o2)
        assert o2["modes"]["tasks_long"] == 0
    # the fan-in-300 graph of the 32-bit-code test: one destination with 300 in-edges
    k = 300
    level_off = [0, 1, 1 + k, 2 + k, 3 + k]
    adj_off = [0, k] + list(range(k + 1, 2 * k + 1)) + [2 * k + 1, 2 * k + 1]
    adj_dst = list(range(1, 1 + k)) + [1 + k] * k + [2 + k]
    adj_w = [0] * k + [i % 2 for i in range(k)] + [0]
    ncol = np.zeros(3 + k, np.int64)
    ncol[1:1 + k] = 1
    g = LevelGraph(level_off, adj_off, adj_dst, adj_w, np.concatenate([[0], np.cumsum(ncol)]), np.arange(k) % 7, [1, 0, 1, 0, 0, 1, 0])
    o = dp_emu.dp_diploid(g, R)
    assert_dip_equal(oracle_dip(oracle_mod, g, R), o)
    assert o["modes"]["tasks_long"] > 0


# ---- row-sharded sweep over several GPUs (dg_dip_create_sharded), emulated rank by rank -------------------------------
@pytest.mark.parametrize("n_ranks,grid,tile_cells", [(2, 4, 256), (3, 2, 64), (4, 3, 300), (8, 1, 128), (2, 4, 16384)])
def test_row_sharded_sweep_matches_oracle(n_ranks, grid, tile_cells, oracle_mod, dp_emu):
    """Wide transitions row-split over n_ranks x grid CTAs with pushed rows and broadcast arrivals, narrow ones
    replicated: every rank ends with the same layers and predecessor codes, and the result is the oracle's."""
    seen = dict(narrow=0, wide=0, wide_tasks=0, pushes=0)
    for seed in range(10):
        rng = np.random.default_rng(900 + seed)
        g = synth.random_level_graph(700 + seed, n_levels=int(rng.integers(3, 30)), max_width=int(rng.integers(2, 14)),
                                     n_colours=int(rng.integers(0, 150)), p_colour=0.5)
        R = int(rng.integers(0, 6))
        o = dp_emu.dp_diploid_sharded(g, R, n_ranks, grid=grid, tile_cells=tile_cells)
        assert_dip_equal(oracle_dip(oracle_mod, g, R), o, checks=False)
        for k in seen:
            seen[k] += o["modes"][k]
    if tile_cells < 1000:
        assert seen["wide"] > 0 and seen["narrow"] > 0 and seen["pushes"] >= seen["wide"]


def test_row_sharded_lane_panel_and_long_destinations(oracle_mod, dp_emu):
    g = synth.lane_panel_graph(11, n_lanes=12, n_blocks=4, rec_per_block=2, p_colour=0.3, n_colours=64)
    for n_ranks in (2, 4):
        o = dp_emu.dp_diploid_sharded(g, 3, n_ranks, grid=4, tile_cells=512)
        assert_dip_equal(oracle_dip(oracle_mod, g, 3), o, checks=False)
        assert o["modes"]["wide"] > 0
    g = funnel_graph(5, lanes=40)
    o = dp_emu.dp_diploid_sharded(g, 2, 2, grid=4, tile_cells=2048)
    assert_dip_equal(oracle_dip(oracle_mod, g, 2), o, checks=False)
