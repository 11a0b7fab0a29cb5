#!/usr/bin/env python
"""Regenerates tests/golden/* by running the reference itself (oracle/_ref/ref_driver, i.e. the
reference's own objects plus read-only dump hooks; see oracle/Makefile) on the bundled test data.

Run in the build container only (needs /root/reference):  python tests/golden/make_goldens.py
Outputs (committed):
  tiny_<name>.dgd            full stage dumps for the toy graphs (a few KB each)
  mhc4_chm13_dipin.npz       the levelized ExpandedGraph the reference hands to its diploid DP for
                             test/MHC_4.gfa.gz + test/CHM13_reads.fq.gz (independent of R)
  mhc4_chm13_hapin.npz       the Kahn-ordered ExpandedGraph handed to the haploid DP
  expected.json              reference results: FASTA md5, DP value, s_het, recombination-edge lists,
                             sha256 of the per-level DP checksums, haploid scores/path digest, log counters
"""
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dipgenie_b200 import dgd  # noqa: E402
from dipgenie_b200.cuda_api import LevelGraph  # noqa: E402

REF_TEST = "/root/reference/test"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLD = os.path.dirname(os.path.abspath(__file__))
TMP = os.path.join(ROOT, "scratch", "goldens")


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def run(name, gfa, reads, extra, keep_dump=False, light=True):
    os.makedirs(TMP, exist_ok=True)
    fa = os.path.join(TMP, name + ".fa")
    dump = os.path.join(TMP, name + ".dgd")
    cmd = [DRIVER, "-g", gfa, "-r", reads, "-o", fa, "-D", dump, "-t", "8"] + (["-L"] if light else []) + extra
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {p.stderr[-2000:]}")
    log = p.stdout + "\n" + p.stderr
    d = dgd.load(dump)
    return fa, d, log, dump


def dip_expected(fa, d, log):
    e = dict(fasta_md5=md5(fa), value=int(d["dip_out.value"][0]), s_het=int(d["dip_out.s_het"][0]),
             p1_edges=d["dip_out.p1_edges"].tolist(), p2_edges=d["dip_out.p2_edges"].tolist(),
             checksum_sha256=hashlib.sha256(d["dip_out.level_checksum"][1:].tobytes()).hexdigest(),
             live_sha256=hashlib.sha256(d["dip_out.level_live"][1:].tobytes()).hexdigest())
    m = re.search(r"recombinations in P1: (\d+), recombinations in P2: (\d+), bp of P1: (\d+), bp of P2: (\d+)", log)
    e.update(r1=int(m.group(1)), r2=int(m.group(2)), bp1=int(m.group(3)), bp2=int(m.group(4)))
    return e


def hap_expected(fa, d, log):
    return dict(fasta_md5=md5(fa), colors_by_r=d["hap_out.colors_by_r"].tolist(), best_r=int(d["hap_out.best_r"][0]),
                path_len=int(len(d["hap_out.path"])), path_sha256=hashlib.sha256(d["hap_out.path"].tobytes()).hexdigest(),
                path_original_sha256=hashlib.sha256(d["hap_out.path_original"].tobytes()).hexdigest())


def counters(log):
    c = {}
    m = re.search(r"spectrum size: (\d+)", log)
    if m:
        c["spectrum"] = int(m.group(1))
    blocks = re.split(r"Number of Anchors", log)
    mins = dict(re.findall(r"^(\S+) : (\d+)$", blocks[0], re.M))
    c["minimizers"] = {k: int(v) for k, v in mins.items()}
    if len(blocks) > 1:
        c["anchors"] = {k: int(v) for k, v in re.findall(r"^(\S+) : (\d+)$", blocks[1], re.M)}
    m = re.search(r"Homozygous: ([\d.]+)%, Heterozygous: ([\d.]+)%, Total kmers: (\d+)", log)
    if m:
        c["hom_pct"], c["het_pct"], c["total_kmers"] = m.group(1), m.group(2), int(m.group(3))
    m = re.search(r"Fitted model: (.*)$", log, re.M)
    if m:
        c["fitted"] = m.group(1).strip()
    return c


def main():
    exp = {"tiny": {}, "mhc4_chm13": {"diploid": {}, "haploid": {}}}
    tiny = [
        ("test_p2_R2_k5_w3", "test.gfa", "read.fa", ["-p", "2", "-R", "2", "-k", "5", "-w", "3"]),
        ("test_p2_R0_k3_w2", "test.gfa", "read.fa", ["-p", "2", "-R", "0", "-k", "3", "-w", "2"]),
        ("test_p2_R1_k3_w2", "test.gfa", "read.fa", ["-p", "2", "-R", "1", "-k", "3", "-w", "2"]),
        ("test_p2_R2_k3_w2", "test.gfa", "read.fa", ["-p", "2", "-R", "2", "-k", "3", "-w", "2"]),
        ("test_p1_R2_k3_w2", "test.gfa", "read.fa", ["-p", "1", "-R", "2", "-k", "3", "-w", "2"]),
        ("test2_p2_R2", "test2.gfa", "read2.fa", ["-p", "2", "-R", "2"]),
        ("test2_p1_R2", "test2.gfa", "read2.fa", ["-p", "1", "-R", "2"]),
    ]
    for name, gfa, reads, extra in tiny:
        fa, d, log, dump = run(name, os.path.join(REF_TEST, gfa), os.path.join(REF_TEST, reads), extra, light=False)
        os.replace(dump, os.path.join(GOLD, f"tiny_{name}.dgd"))
        e = dict(args=extra, gfa=gfa, reads=reads, fasta=open(fa).read(), counters=counters(log))
        e.update(dip_expected(fa, d, log) if "dip_out.value" in d else hap_expected(fa, d, log))
        exp["tiny"][name] = e
        print("tiny", name, e.get("value", e.get("best_r")))

    gfa, reads = os.path.join(REF_TEST, "MHC_4.gfa.gz"), os.path.join(REF_TEST, "CHM13_reads.fq.gz")
    for R in (18, 0, 6, 36):
        fa, d, log, dump = run(f"mhc_p2_R{R}", gfa, reads, ["-p", "2", "-R", str(R)])
        e = dip_expected(fa, d, log)
        exp["mhc4_chm13"]["diploid"][str(R)] = e
        if R == 18:
            exp["mhc4_chm13"]["counters"] = counters(log)
            LevelGraph.from_dgd(d).to_npz(os.path.join(GOLD, "mhc4_chm13_dipin.npz"),
                                          checksum_every64=d["dip_out.level_checksum"][::64].copy())
        os.remove(dump)
        print("mhc diploid R", R, e["value"], e["fasta_md5"])
    fa, d, log, dump = run("mhc_p1", gfa, reads, ["-p", "1"])
    exp["mhc4_chm13"]["haploid"]["18"] = hap_expected(fa, d, log)
    hg = LevelGraph(np.arange(2, dtype=np.int32), d["hap_in.adj.off"], d["hap_in.adj.dst"], d["hap_in.adj.w"],
                    d["hap_in.color.off"], d["hap_in.color.val"], np.zeros(1, np.uint8))
    hg.to_npz(os.path.join(GOLD, "mhc4_chm13_hapin.npz"))
    os.remove(dump)
    print("mhc haploid", exp["mhc4_chm13"]["haploid"]["18"]["fasta_md5"])
    json.dump(exp, open(os.path.join(GOLD, "expected.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
