"""CPU-side checks of the level-program engine (v4) of the diploid DP: the host planning (dp_plan4.cpp), the
descriptor functions the device builder runs (dp_prog.h) and the program interpreter's semantics
(tests/emu/dp_emu4.cpp mirrors dip_sweep4_kernel) against the reference's goldens and the oracle — value, s_het, both
recombination-edge lists and the per-level checksums of every live cell."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD, assert_dip_equal, oracle_dip
from dipgenie_b200 import dgd, synth
from dipgenie_b200.cuda_api import LevelGraph

TINY_DIP = ["test_p2_R2_k5_w3", "test_p2_R0_k3_w2", "test_p2_R1_k3_w2", "test_p2_R2_k3_w2", "test2_p2_R2"]
# (slog, kn, slot_bytes, grid, rc): default; tiny tiles (most levels in HBM, several CTAs); tiny slots (programs read in
# place); other layer chunkings
SHAPES = [None, (4, 4, 8192, 3, 10), (10, 32, 64, 1, 4), (6, 8, 256, 2, 7), (9, 22, 8192, 1, 5)]


@pytest.mark.parametrize("name", TINY_DIP)
def test_program_matches_reference_tiny(name, dp_emu4, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = LevelGraph.from_dgd(d)
    R = int(d["dip_in.R"][0])
    for shape in SHAPES:
        o = dp_emu4.dp_diploid(g, R, shape=shape)
        e = expected["tiny"][name]
        assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
        assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
        assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
        assert np.array_equal(o["checksum"][1:], d["dip_out.level_checksum"][1:])
        assert np.array_equal(o["live"][1:], d["dip_out.level_live"][1:])


def test_program_matches_reference_mhc(dp_emu4, expected):
    """MHC_4 + CHM13 reads, R = 6: every one of the 120 362 layers' checksums equals the reference's."""
    g, _ = LevelGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"))
    o = dp_emu4.dp_diploid(g, 6)
    e = expected["mhc4_chm13"]["diploid"]["6"]
    assert o["value"] == e["value"] and o["s_het"] == e["s_het"]
    assert o["p1_edges"].ravel().tolist() == e["p1_edges"]
    assert o["p2_edges"].ravel().tolist() == e["p2_edges"]
    assert hashlib.sha256(o["checksum"][1:].tobytes()).hexdigest() == e["checksum_sha256"]
    m = o["modes"]
    assert m["smem"] > 100000 and m["all_ctas"] > 100 and m["compact"] > 90000 and m["big_cells"] > 0
    assert m["skipped"] > 5000 and m["relocations"] > 10 and m["cells_written"] < 0.6 * m["cells_total"]
    print(m)


@pytest.mark.parametrize("seed", range(12))
def test_program_matches_oracle_random(seed, oracle_mod, dp_emu4):
    rng = np.random.default_rng(seed)
    g = synth.random_level_graph(seed, n_levels=int(rng.integers(3, 40)), max_width=int(rng.integers(2, 14)),
                                 p_weight1=0.3, p_colour=0.4, n_colours=int(rng.integers(1, 200)))
    R = int(rng.integers(0, 7))
    ref = oracle_dip(oracle_mod, g, R)
    for shape in SHAPES:
        assert_dip_equal(ref, dp_emu4.dp_diploid(g, R, shape=shape))


def test_program_lane_panels(oracle_mod, dp_emu4):
    for seed, lanes, R in ((1, 6, 3), (2, 12, 5), (3, 40, 4)):
        g = synth.lane_panel_graph(seed, n_lanes=lanes, n_blocks=5, rec_per_block=2, p_colour=0.3, n_colours=96)
        ref = oracle_dip(oracle_mod, g, R)
        for shape in (None, (8, 16, 8192, 4, 10)):
            o = dp_emu4.dp_diploid(g, R, shape=shape)
            assert o is not None
            assert_dip_equal(ref, o)
            if lanes * lanes > 1024:     # recombination x recombination cells: ordinals beyond the packed key (dp_prog.h: PROG_KEY_CAND)
                assert o["modes"]["max_cand"] > 1024


def test_planner_rejects_what_the_kernels_cannot_take(dp_emu4, dp_emu):
    """C-ABI input validation (dp_prep.cpp: build_dip_plan): weights above 1, and parallel edges of differing weight —
    the reference resolves their ties by thread timing (approximator.cpp:627-701), so there is nothing to be exact with."""
    ok = LevelGraph([0, 1, 3, 4], [0, 2, 3, 4, 4], [1, 2, 3, 3], [0, 1, 0, 0], [0, 0, 0, 0, 0], [], [0])
    assert dp_emu4.dp_diploid(ok, 2) is not None
    heavy = LevelGraph([0, 1, 3, 4], [0, 2, 3, 4, 4], [1, 2, 3, 3], [0, 2, 0, 0], [0, 0, 0, 0, 0], [], [0])
    mixed = LevelGraph([0, 1, 2, 3], [0, 2, 3, 3], [1, 1, 2], [0, 1, 0], [0, 0, 0, 0], [], [0])
    same = LevelGraph([0, 1, 2, 3], [0, 2, 3, 3], [1, 1, 2], [1, 1, 0], [0, 0, 0, 0], [], [0])     # duplicates of equal weight are fine
    assert dp_emu4.dp_diploid(same, 2) is not None
    for bad in (heavy, mixed):
        with pytest.raises(RuntimeError):
            dp_emu4.dp_diploid(bad, 2)
        with pytest.raises(RuntimeError):
            dp_emu.dp_diploid(bad, 2)
