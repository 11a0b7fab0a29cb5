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


def test_front_end_builds_the_reference_dp_input_on_a_90_walk_panel(tmp_path):
    """The host glue (GFA reader, panel columns, anchor filter and order, classifier fit, expansion, Kahn order, levelization)
    with the oracle standing in for the device stages, on the config-4 panel at 1/16 of the backbone: the DP input it dumps
    (DG_DUMP_DIPIN) is array-for-array the one the unmodified reference built for the same files (digest recorded from
    oracle/_ref/ref_driver's dump)."""
    import subprocess
    import make_config4
    from dipgenie_b200 import _build, dgd
    import oracle
    oracle.build(); _build.build_host()
    exe = os.path.join(ROOT, "tests", "host", "host_check")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(ROOT, "tests", "host", "host_check.cpp"),
                               "-L" + _build.PKG, "-ldipgenie_host", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                               "-Wl,-rpath," + _build.PKG, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lz", "-lm"])
    e = json.load(open(os.path.join(GOLD, "e2e_expected.json")))["c4_h90_s16_p2_R18"]
    gfa, fa = make_config4.make(str(tmp_path), scale=0.0625)
    dump = str(tmp_path / "dipin.dgd")
    env = dict(os.environ, DG_DUMP_DIPIN=dump, DG_DUMP_ONLY="1")
    p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", str(tmp_path / "o.fa"), "-t", "8", "-p", "2", "-R", "18"], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    a = dgd.load(dump)
    h = hashlib.sha256()
    for name, dt in (("level_off", np.int64), ("adj_off", np.int64), ("adj_dst", np.int32), ("adj_w", np.uint8), ("col_off", np.int64),
                     ("col_val", np.int32), ("colour_is_hom", np.uint8)):
        h.update(np.ascontiguousarray(np.asarray(a[name]).astype(dt)).tobytes())
    assert h.hexdigest() == e["dipin_sha256"]
