"""CPU-side checks of the haploid DP path (no GPU needed): the oracle (oracle/dp_haploid.c) against the
reference's own results in tests/golden/ (distinct colours per r, the traceback path of best_r), on the
toy graphs and on the full MHC_4 + CHM13 input (n = 499 223 vertices, R = 18)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLD
from dipgenie_b200 import dgd
from dipgenie_b200.cuda_api import HapGraph

TINY_HAP = ["test_p1_R2_k3_w2", "test2_p1_R2"]


@pytest.mark.parametrize("name", TINY_HAP)
def test_oracle_matches_reference_tiny(name, oracle_mod, expected):
    d = dgd.load(os.path.join(GOLD, f"tiny_{name}.dgd"))
    g = HapGraph.from_dgd(d)
    R = int(d["hap_in.R"][0])
    o = oracle_mod.dp_haploid(g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.n_colours, R)
    e = expected["tiny"][name]
    assert o["colours_by_r"].tolist() == e["colors_by_r"] == d["hap_out.colors_by_r"].tolist()
    best = int(d["hap_out.best_r"][0])
    assert np.array_equal(o["paths"][best], d["hap_out.path"])


def test_oracle_matches_reference_mhc(oracle_mod, expected):
    g = HapGraph.from_npz(os.path.join(GOLD, "mhc4_chm13_hapin.npz"))
    e = expected["mhc4_chm13"]["haploid"]["18"]
    o = oracle_mod.dp_haploid(g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.n_colours, 18)
    assert o["colours_by_r"].tolist() == e["colors_by_r"]
    p = o["paths"][e["best_r"]].astype(np.int32)
    assert len(p) == e["path_len"]
    assert hashlib.sha256(p.tobytes()).hexdigest() == e["path_sha256"]
