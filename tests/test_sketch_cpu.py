"""CPU-side pin of the sketch oracle (oracle/sketch.c) against the reference's own Solver::index_kmers /
compute_hashes outputs (tests/golden/sketch_*.npz, produced by tests/golden/make_sketch_goldens.py):
toy inputs in full, the MHC_4 panel + CHM13 reads through sha256 digests of every output array."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD

TINY = ["test_k3_w2", "test_k5_w3", "test2_k31_w25"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_oracle(oracle_mod, z):
    k, w = int(z["k"]), int(z["w"])
    sp, cnt, roff, rval = oracle_mod.sketch_reads(z["read_bases"], z["read_off"], k, w, per_read=True)
    ix = oracle_mod.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, sp,
                                want_all=True)
    return sp, cnt, roff, rval, ix


@pytest.mark.parametrize("name", TINY)
def test_oracle_matches_reference_tiny(name, oracle_mod):
    z = dict(np.load(os.path.join(GOLD, f"sketch_{name}.npz")))   # materialise once: NpzFile decompresses on every access
    sp, cnt, roff, rval, ix = run_oracle(oracle_mod, z)
    assert np.array_equal(roff.astype(np.int64), z["read_hashes_off"]) and np.array_equal(rval, z["read_hashes_val"])
    assert np.array_equal(sp, np.unique(z["read_hashes_val"]))
    nw = len(z["walk_off"]) - 1
    spset = set(sp.tolist())
    for h in range(nw):
        ref_hash = z[f"index{h}_hash"]
        assert np.array_equal(ix["all_hash"][ix["all_off"][h]:ix["all_off"][h + 1]], ref_hash)
        assert ix["n_minimizers"][h] == len(ref_hash)
        # the join keeps, in walk order, the index entries whose hash is in the read spectrum (solver.cpp:415-446)
        keep = [i for i, x in enumerate(ref_hash.tolist()) if x in spset]
        lo, hi = int(ix["hit_off"][h]), int(ix["hit_off"][h + 1])
        assert hi - lo == len(keep)
        assert np.array_equal(sp[ix["hit_sid"][lo:hi]], ref_hash[keep])
        voff, vval = z[f"index{h}_vtx_off"], z[f"index{h}_vtx_val"]
        for t, i in enumerate(keep):
            a, b = int(ix["hit_vtx_off"][lo + t]), int(ix["hit_vtx_off"][lo + t + 1])
            assert np.array_equal(ix["hit_vtx"][a:b], vval[voff[i]:voff[i + 1]])


def test_oracle_matches_reference_mhc(oracle_mod):
    z = dict(np.load(os.path.join(GOLD, "sketch_mhc4_chm13.npz")))   # materialise once: NpzFile decompresses on every access
    e = json.load(open(os.path.join(GOLD, "sketch_expected.json")))
    k, w = int(z["k"]), int(z["w"])
    sp, cnt, roff, rval = oracle_mod.sketch_reads(z["read_bases"], z["read_off"], k, w, per_read=True)
    assert sha(roff.astype(np.int64)) == e["reads"]["off_sha256"] and sha(rval) == e["reads"]["val_sha256"]
    assert len(sp) == e["spectrum"]["n"] and sha(sp) == e["spectrum"]["sha256"]
    assert int(cnt.sum()) == len(rval)
    # full join: every walk minimizer is looked up in a spectrum that contains all of them, so hits == index
    ix = oracle_mod.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, sp,
                                want_all=True)
    for h, ew in enumerate(e["walks"]):
        allh = ix["all_hash"][ix["all_off"][h]:ix["all_off"][h + 1]]
        assert len(allh) == ew["n"] and sha(allh) == ew["hash_sha256"]
    full = np.unique(ix["all_hash"])
    ix2 = oracle_mod.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, full)
    for h, ew in enumerate(e["walks"]):
        lo, hi = int(ix2["hit_off"][h]), int(ix2["hit_off"][h + 1])
        assert hi - lo == ew["n"]
        voff = ix2["hit_vtx_off"][lo:hi + 1].astype(np.int64)
        assert sha(voff - voff[0]) == ew["vtx_off_sha256"]
        assert sha(ix2["hit_vtx"][int(voff[0]):int(voff[-1])].astype(np.int32)) == ew["vtx_val_sha256"]
