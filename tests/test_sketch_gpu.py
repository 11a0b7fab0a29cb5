"""GPU parity tests of the minimizer sketch / read spectrum / walk-index join (through the C ABI):
CUDA vs the oracle on seeded inputs (raw minimizer lists, spectrum, read counts, hits and covered
vertices — all bit-exact), vs the reference's committed outputs on the toy inputs, and on the full
MHC_4 panel + CHM13 reads vs the reference's sha256 digests."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from dipgenie_b200.cuda_api import Context

pytestmark = pytest.mark.gpu

TINY = ["test_k3_w2", "test_k5_w3", "test2_k31_w25"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def random_seqs(rng, n_seq, max_len, alphabet, p_lower=0.0):
    seqs = []
    for _ in range(n_seq):
        n = int(rng.integers(0, max_len + 1))
        s = rng.choice(np.frombuffer(alphabet, np.uint8), n)
        if p_lower:
            low = rng.random(n) < p_lower
            s = np.where(low, s | 0x20, s).astype(np.uint8)
        seqs.append(s.astype(np.uint8))
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    bases = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    return bases.astype(np.uint8), off


def oracle_raw(oracle_mod, bases, off, k, w):
    """Raw per-sequence minimizer hashes via the walk oracle: every sequence is a one-segment walk."""
    n = len(off) - 1
    ix = oracle_mod.index_walks(bases, off, np.arange(n, dtype=np.int32), np.arange(n + 1, dtype=np.uint64),
                                np.arange(max(n, 1), dtype=np.int32), k, w, np.zeros(0, np.uint64), want_all=True)
    return ix["all_off"], ix["all_hash"]


@pytest.mark.parametrize("case", [
    dict(seed=1, n_seq=40, max_len=300, alphabet=b"ACGT", k=31, w=25),
    dict(seed=2, n_seq=30, max_len=200, alphabet=b"ACGT", k=5, w=3, p_lower=0.3),
    dict(seed=3, n_seq=25, max_len=400, alphabet=b"ACGTN", k=31, w=25),           # 4 bits/symbol, two-word keys
    dict(seed=4, n_seq=25, max_len=300, alphabet=b"ACGTNRYKM*-", k=21, w=11, p_lower=0.2),
    dict(seed=5, n_seq=10, max_len=9000, alphabet=b"ACGT", k=31, w=25),           # sequences spanning several tiles
    dict(seed=6, n_seq=60, max_len=80, alphabet=b"AC", k=7, w=1),                 # low complexity: many equal k-mers, w=1
    dict(seed=7, n_seq=8, max_len=5000, alphabet=b"ACGTacgtnN", k=16, w=50),
    dict(seed=8, n_seq=12, max_len=700, alphabet=bytes(range(33, 127)), k=12, w=9),   # 8 bits/symbol
])
def test_minimizers_match_oracle(case, ctx, oracle_mod):
    rng = np.random.default_rng(case["seed"])
    bases, off = random_seqs(rng, case["n_seq"], case["max_len"], case["alphabet"], case.get("p_lower", 0.0))
    k, w = case["k"], case["w"]
    cnt, hashes, starts = ctx.sketch_minimizers(bases, off, k, w)
    aoff, ahash = oracle_raw(oracle_mod, bases, off, k, w)
    assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64), aoff)
    assert np.array_equal(hashes, ahash)
    # starts are valid k-mer starts and the hash is the hash of that canonical k-mer
    lens = np.diff(off.astype(np.int64))
    seq_of = np.repeat(np.arange(len(cnt)), cnt.astype(np.int64))
    assert np.all(starts.astype(np.int64) + k <= lens[seq_of])
    sp, rc = ctx.sketch_reads(bases, off, k, w)
    osp, orc = oracle_mod.sketch_reads(bases, off, k, w)
    assert np.array_equal(sp, osp) and np.array_equal(rc, orc)


def test_empty_and_short_inputs(ctx):
    off = np.array([0, 0, 10, 10], np.uint64)
    bases = np.frombuffer(b"ACGTACGTAC", np.uint8)
    cnt, h, s = ctx.sketch_minimizers(bases, off, 31, 25)        # everything shorter than w+k-1 (:372)
    assert cnt.tolist() == [0, 0, 0] and len(h) == 0
    sp, rc = ctx.sketch_reads(bases, off, 31, 25)
    assert len(sp) == 0 and len(rc) == 0
    cnt, h, s = ctx.sketch_minimizers(np.zeros(0, np.uint8), np.array([0], np.uint64), 31, 25)
    assert len(h) == 0


def random_panel(rng, n_seg, n_walks, alphabet=b"ACGT"):
    seg = [rng.choice(np.frombuffer(alphabet, np.uint8), int(rng.integers(0, 60))).astype(np.uint8) for _ in range(n_seg)]
    seg_off = np.concatenate([[0], np.cumsum([len(s) for s in seg])]).astype(np.uint64)
    walks = []
    for _ in range(n_walks):
        keep = np.sort(rng.choice(n_seg, int(rng.integers(1, n_seg + 1)), replace=False))
        walks.append(keep.astype(np.int32))
    walk_off = np.concatenate([[0], np.cumsum([len(x) for x in walks])]).astype(np.uint64)
    return np.concatenate(seg), seg_off, np.concatenate(walks), walk_off, rng.permutation(n_seg).astype(np.int32)


@pytest.mark.parametrize("seed", range(6))
def test_index_walks_match_oracle(seed, ctx, oracle_mod):
    rng = np.random.default_rng(500 + seed)
    k, w = (31, 25) if seed % 2 == 0 else (9, 4)
    segb, sego, wv, wo, tom = random_panel(rng, int(rng.integers(20, 400)), int(rng.integers(1, 7)),
                                           b"ACGT" if seed < 4 else b"ACGTN")
    # reads: substrings of walk 0 so that a good share of minimizers hit
    w0 = np.concatenate([segb[int(sego[v]):int(sego[v + 1])] for v in wv[int(wo[0]):int(wo[1])]] + [np.zeros(0, np.uint8)])
    reads = [w0[a:a + 150] for a in rng.integers(0, max(1, len(w0) - 150), 30)]
    roff = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    rb = np.concatenate(reads).astype(np.uint8)
    sp, _ = ctx.sketch_reads(rb, roff, k, w)
    a = ctx.index_walks(segb, sego, wv, wo, tom, k, w, sp)
    b = oracle_mod.index_walks(segb, sego, wv, wo, tom, k, w, sp)
    for key in ("n_minimizers", "hit_off", "hit_sid", "hit_vtx_off", "hit_vtx"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("name", TINY)
def test_cuda_matches_reference_tiny(name, ctx):
    z = dict(np.load(os.path.join(GOLD, f"sketch_{name}.npz")))   # materialise once: NpzFile decompresses on every access
    k, w = int(z["k"]), int(z["w"])
    sp, rc = ctx.sketch_reads(z["read_bases"], z["read_off"], k, w)
    assert np.array_equal(sp, np.unique(z["read_hashes_val"]))
    nw = len(z["walk_off"]) - 1
    full = np.unique(np.concatenate([z[f"index{h}_hash"] for h in range(nw)]))
    ix = ctx.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, full)
    for h in range(nw):
        lo, hi = int(ix["hit_off"][h]), int(ix["hit_off"][h + 1])
        assert np.array_equal(full[ix["hit_sid"][lo:hi]], z[f"index{h}_hash"])
        voff = ix["hit_vtx_off"][lo:hi + 1].astype(np.int64)
        assert np.array_equal(voff - voff[0], z[f"index{h}_vtx_off"])
        assert np.array_equal(ix["hit_vtx"][int(voff[0]):int(voff[-1])], z[f"index{h}_vtx_val"])


def test_cuda_matches_reference_mhc(ctx):
    """Full size: 16 401 reads, 5 walks x ~5 Mbp; every output array against the reference's digests."""
    z = dict(np.load(os.path.join(GOLD, "sketch_mhc4_chm13.npz")))   # materialise once: NpzFile decompresses on every access
    e = json.load(open(os.path.join(GOLD, "sketch_expected.json")))
    k, w = int(z["k"]), int(z["w"])
    sp, rc = ctx.sketch_reads(z["read_bases"], z["read_off"], k, w)
    assert len(sp) == e["spectrum"]["n"] and sha(sp) == e["spectrum"]["sha256"]
    cnt, hashes, starts = ctx.sketch_minimizers(z["read_bases"], z["read_off"], k, w)
    # per-read sets (std::set iteration order) rebuilt from the raw lists
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    sets = [np.unique(hashes[off[i]:off[i + 1]]) for i in range(len(cnt))]
    soff = np.concatenate([[0], np.cumsum([len(x) for x in sets])]).astype(np.int64)
    assert sha(soff) == e["reads"]["off_sha256"] and sha(np.concatenate(sets)) == e["reads"]["val_sha256"]
    assert int(rc.sum()) == int(soff[-1])
    # walks: join against a spectrum that contains every walk minimizer -> hits == the whole index
    nw = len(z["walk_off"]) - 1
    wseq = [np.concatenate([z["seg_bases"][int(z["seg_off"][v]):int(z["seg_off"][v + 1])]
                            for v in z["walk_vtx"][int(z["walk_off"][h]):int(z["walk_off"][h + 1])]]) for h in range(nw)]
    woff = np.concatenate([[0], np.cumsum([len(x) for x in wseq])]).astype(np.uint64)
    wc, wh, ws = ctx.sketch_minimizers(np.concatenate(wseq), woff, k, w)
    assert wc.tolist() == [x["n"] for x in e["walks"]]
    full = np.unique(wh)
    ix = ctx.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, full)
    assert ix["n_minimizers"].tolist() == [x["n"] for x in e["walks"]]
    for h, ew in enumerate(e["walks"]):
        lo, hi = int(ix["hit_off"][h]), int(ix["hit_off"][h + 1])
        assert hi - lo == ew["n"]
        assert sha(full[ix["hit_sid"][lo:hi]]) == ew["hash_sha256"]
        voff = ix["hit_vtx_off"][lo:hi + 1].astype(np.int64)
        assert sha(voff - voff[0]) == ew["vtx_off_sha256"]
        assert sha(ix["hit_vtx"][int(voff[0]):int(voff[-1])].astype(np.int32)) == ew["vtx_val_sha256"]
    # and the real join: hits against the read spectrum are the index entries whose hash is in Sp_R
    ix2 = ctx.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, sp)
    woff_i = np.concatenate([[0], np.cumsum(wc)]).astype(np.int64)
    for h in range(nw):
        hh = wh[woff_i[h]:woff_i[h + 1]]
        keep = np.isin(hh, sp)
        lo, hi = int(ix2["hit_off"][h]), int(ix2["hit_off"][h + 1])
        assert hi - lo == int(keep.sum())
        assert np.array_equal(sp[ix2["hit_sid"][lo:hi]], hh[keep])
