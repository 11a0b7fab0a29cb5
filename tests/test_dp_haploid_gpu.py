"""GPU parity tests of the haploid DP (through the C ABI, via ctypes): CUDA path vs the oracle on seeded
random DAGs (all R+1 traceback paths and distinct-colour counts, bit-exact), vs the reference's committed
goldens on the toy graphs, and at full size on the bundled MHC_4 + CHM13 input."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD
from dipgenie_b200 import dgd, synth
from dipgenie_b200.cuda_api import Context, HapGraph

pytestmark = pytest.mark.gpu

TINY_HAP = ["test_p1_R2_k3_w2", "test2_p1_R2"]


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def oracle_hap(oracle_mod, g, R):
    return oracle_mod.dp_haploid(g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.n_colours, R)


def assert_hap_equal(a, b):
    assert a["colours_by_r"].tolist() == b["colours_by_r"].tolist()
    assert len(a["paths"]) == len(b["paths"])
    for r, (x, y) in enumerate(zip(a["paths"], b["paths"])):
        assert np.array_equal(x, y), f"path of layer {r} differs"


@pytest.mark.parametrize("name", TINY_HAP)
def test_cuda_matches_reference_tiny(name, ctx, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = HapGraph.from_dgd(d)
    R = int(d["hap_in.R"][0])
    o = ctx.dp_haploid(g, R)
    assert o["colours_by_r"].tolist() == expected["tiny"][name]["colors_by_r"]
    assert np.array_equal(o["paths"][int(d["hap_out.best_r"][0])], d["hap_out.path"])


@pytest.mark.parametrize("seed", range(16))
def test_cuda_matches_oracle_random(seed, ctx, oracle_mod):
    rng = np.random.default_rng(7000 + seed)
    g = synth.random_kahn_graph(seed, n=int(rng.integers(2, 600)), max_out=int(rng.integers(1, 6)),
                                p_weight1=float(rng.random() * 0.7), p_colour=float(rng.random()),
                                n_colours=int(rng.integers(1, 300)), max_span=int(rng.integers(1, 40)))
    R = int(rng.integers(0, 10))
    assert_hap_equal(oracle_hap(oracle_mod, g, R), ctx.dp_haploid(g, R))


def test_cuda_edge_cases(ctx, oracle_mod):
    one = HapGraph([0, 0], [], [], [0, 0], [], n_colours=1)          # a single vertex: every path is [0]
    assert_hap_equal(oracle_hap(oracle_mod, one, 2), ctx.dp_haploid(one, 2))
    # no colours anywhere: nothing beats the initial 0, every traceback stops at the sink (:50-52)
    g = synth.random_kahn_graph(3, n=50, p_colour=0.0)
    o = ctx.dp_haploid(g, 3)
    assert_hap_equal(oracle_hap(oracle_mod, g, 3), o)
    assert all(len(p) == 1 for p in o["paths"])
    # malformed input is an error, not a hang
    bad = HapGraph([0, 1, 2], [1, 0], [0, 0], [0, 0, 0], [], n_colours=1)
    with pytest.raises(Exception):
        ctx.dp_haploid(bad, 1)


def test_cuda_matches_reference_mhc(ctx, expected):
    g = HapGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_hapin.npz"))
    e = expected["mhc4_chm13"]["haploid"]["18"]
    p = ctx.hap_create(g, 18)
    try:
        p.run()
        res = p.result()
        assert res["colours_by_r"].tolist() == e["colors_by_r"]
        path = p.path(e["best_r"], int(res["path_len"][e["best_r"]])).astype(np.int32)
        st = p.stats()
    finally:
        p.close()
    assert len(path) == e["path_len"]
    assert hashlib.sha256(path.tobytes()).hexdigest() == e["path_sha256"]
    assert st["cell_updates"] == 19 * 924889
