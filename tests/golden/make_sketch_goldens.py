#!/usr/bin/env python
"""Regenerates the sketch-stage fixtures by running the reference itself (oracle/_ref/ref_driver -S: the
reference's own Solver::index_kmers / compute_hashes / compute_and_classify_anchors, dumped read-only).

Run in the build container only (needs /root/reference):  python tests/golden/make_sketch_goldens.py
Outputs (committed):
  sketch_<name>.npz       toy inputs (segment sequences, walks, top_order_map, reads) together with the
                          reference's per-walk minimizer index and per-read hash sets, in full
  sketch_mhc4_chm13.npz   the same inputs for test/MHC_4.gfa.gz + test/CHM13_reads.fq.gz
  sketch_expected.json    sha256 digests / counts of the reference's outputs on the MHC input
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dipgenie_b200 import dgd  # noqa: E402

REF_TEST = "/root/reference/test"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLD = os.path.dirname(os.path.abspath(__file__))
TMP = os.path.join(ROOT, "scratch", "goldens")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(name, gfa, reads, k, w):
    os.makedirs(TMP, exist_ok=True)
    dump = os.path.join(TMP, f"sketch_{name}.dgd")
    cmd = [DRIVER, "-g", os.path.join(REF_TEST, gfa), "-r", os.path.join(REF_TEST, reads), "-o", "/dev/null", "-D", dump,
           "-t", "8", "-S", "-k", str(k), "-w", str(w)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {p.stderr[-2000:]}")
    d = dgd.load(dump)
    os.remove(dump)
    return d


def inputs(d):
    return dict(seg_bases=d["panel.node_seq"], seg_off=d["panel.node_seq_off"].astype(np.uint64),
                walk_vtx=d["panel.paths.val"].astype(np.int32), walk_off=d["panel.paths.off"].astype(np.uint64),
                top_order_map=d["panel.top_order_map"].astype(np.int32), read_bases=d["reads.bases"],
                read_off=d["reads.off"].astype(np.uint64), k=np.int32(0), w=np.int32(0))


def main():
    tiny = [("test_k3_w2", "test.gfa", "read.fa", 3, 2), ("test_k5_w3", "test.gfa", "read.fa", 5, 3),
            ("test2_k31_w25", "test2.gfa", "read2.fa", 31, 25)]
    for name, gfa, reads, k, w in tiny:
        d = run(name, gfa, reads, k, w)
        a = inputs(d)
        a["k"], a["w"] = np.int32(k), np.int32(w)
        nw = int(d["panel.num_walks"][0])
        for h in range(nw):
            a[f"index{h}_hash"] = d[f"index.{h}.hash"]
            a[f"index{h}_vtx_off"] = d[f"index.{h}.vtx.off"]
            a[f"index{h}_vtx_val"] = d[f"index.{h}.vtx.val"]
        a["read_hashes_off"], a["read_hashes_val"] = d["read_hashes.off"], d["read_hashes.val"]
        np.savez_compressed(os.path.join(GOLD, f"sketch_{name}.npz"), **a)
        print("tiny", name, [len(d[f"index.{h}.hash"]) for h in range(nw)], len(d["read_hashes.val"]))
    d = run("mhc", "MHC_4.gfa.gz", "CHM13_reads.fq.gz", 31, 25)
    a = inputs(d)
    a["k"], a["w"] = np.int32(31), np.int32(25)
    np.savez_compressed(os.path.join(GOLD, "sketch_mhc4_chm13.npz"), **a)
    nw = int(d["panel.num_walks"][0])
    exp = {"walks": [], "k": 31, "w": 25}
    for h in range(nw):
        exp["walks"].append(dict(n=int(len(d[f"index.{h}.hash"])), hash_sha256=sha(d[f"index.{h}.hash"]),
                                 vtx_off_sha256=sha(d[f"index.{h}.vtx.off"].astype(np.int64)),
                                 vtx_val_sha256=sha(d[f"index.{h}.vtx.val"].astype(np.int32))))
    exp["reads"] = dict(n=int(d["reads.n"][0]), off_sha256=sha(d["read_hashes.off"].astype(np.int64)),
                        val_sha256=sha(d["read_hashes.val"]))
    sp = np.unique(d["read_hashes.val"])
    exp["spectrum"] = dict(n=int(len(sp)), sha256=sha(sp), count_sp_r=int(d["anchors.count_sp_r"][0]))
    json.dump(exp, open(os.path.join(GOLD, "sketch_expected.json"), "w"), indent=1, sort_keys=True)
    print("mhc", [x["n"] for x in exp["walks"]], exp["spectrum"])


if __name__ == "__main__":
    main()
