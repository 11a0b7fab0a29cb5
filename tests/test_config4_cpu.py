"""CPU checks around the config-4 tooling: the seeded generator is deterministic (the GPU CLI test and bench.py
regenerate the panel on the GPU box and rely on the files being the ones the reference golden was recorded on), the
replicated-panel fixture really replicates, and the traceback's checkpoint chooser keeps its contract."""
import hashlib
import json
import os
import sys

import numpy as np

from conftest import GOLD, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_config4_generator_is_deterministic(tmp_path):
    import make_config4
    e = json.load(open(os.path.join(GOLD, "e2e_expected.json")))["c4_h90_s16_p2_R18"]
    gfa, fa = make_config4.make(str(tmp_path), scale=0.0625)
    assert md5(gfa) == e["gfa_md5"] and md5(fa) == e["reads_md5"]
    walks = [ln for ln in open(gfa, "rb") if ln.startswith(b"W\t")]
    assert len(walks) == 90


def test_replicated_panel_fixture(tmp_path):
    from dipgenie_b200 import fixtures
    g1, _ = fixtures.materialize_mhc(GOLD, str(tmp_path))
    g4, _ = fixtures.materialize_mhc_replicated(GOLD, str(tmp_path), 4)
    w1 = [ln.split(b"\t") for ln in open(g1, "rb") if ln.startswith(b"W\t")]
    w4 = [ln.split(b"\t") for ln in open(g4, "rb") if ln.startswith(b"W\t")]
    assert len(w4) == 4 * len(w1) == 20
    for r in range(4):
        for h, w in enumerate(w1):
            assert w4[r * len(w1) + h][6] == w[6]                       # the same steps
    assert len({(w[1], w[2]) for w in w4}) == 20                        # distinct (sample, haplotype) names
    s1 = [ln for ln in open(g1, "rb") if ln[:1] in b"SL"]
    s4 = [ln for ln in open(g4, "rb") if ln[:1] in b"SL"]
    assert s1 == s4


def test_checkpoint_chooser_contract(dp_emu4, dp_emu):
    """choose_checkpoints (dp_prep.cpp) through the task-stream emulator's checkpointed traceback (tests/emu/dp_emu.cpp runs the
    same anc / hop / seg scheme serially) and the level-program emulator: a long graph with recombination blocks still traces
    to the oracle's edge lists."""
    import oracle
    from dipgenie_b200 import synth
    g = synth.lane_panel_graph(77, n_lanes=6, n_blocks=40, rec_per_block=2, p_colour=0.2, n_colours=64)
    o3 = dp_emu.dp_diploid(g, 3)
    o = dp_emu4.dp_diploid(g, 3)
    ref = oracle.dp_diploid(g.level_off, g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.colour_is_hom, 3)
    for x in (o, o3):
        assert x["value"] == ref["value"] and np.array_equal(x["p1_edges"], ref["p1_edges"]) and np.array_equal(x["p2_edges"], ref["p2_edges"])
