"""Materialises GFA / FASTA input files from the compact panel fixtures under tests/golden/ (segment sequences,
walks, links, reads), so that the CLI can be run end to end on a box that does not have the reference's test/
directory (the GPU box).  The files are equivalent inputs, not copies: segments are renamed 1..n in id order,
links keep their order, optional tags are dropped; tests/golden/make_e2e_inputs.py checks with the reference
binary that they give byte-identical FASTA output."""
from __future__ import annotations

import os

import numpy as np


def write_gfa(path: str, seg_bases, seg_off, walk_vtx, walk_off, link_src, link_dst, walk_sample, walk_hap) -> None:
    seg_bases = np.asarray(seg_bases, np.uint8)
    raw = seg_bases.tobytes()
    n = len(seg_off) - 1
    with open(path, "wb") as f:
        f.write(b"H\tVN:Z:1.1\n")
        for v in range(n):
            f.write(b"S\t%d\t" % (v + 1) + raw[int(seg_off[v]):int(seg_off[v + 1])] + b"\n")
        for a, b in zip(np.asarray(link_src).tolist(), np.asarray(link_dst).tolist()):
            f.write(b"L\t%d\t+\t%d\t+\t0M\n" % (a + 1, b + 1))
        for h in range(len(walk_off) - 1):
            vs = np.asarray(walk_vtx[int(walk_off[h]):int(walk_off[h + 1])]) + 1
            steps = b"".join(b">%d" % x for x in vs.tolist())
            ln = int(sum(int(seg_off[x]) - int(seg_off[x - 1]) for x in vs.tolist()))
            f.write(b"W\t" + str(walk_sample[h]).encode() + b"\t%d\tchr\t0\t%d\t" % (int(walk_hap[h]), ln) + steps + b"\n")


def write_fasta(path: str, bases, off, prefix: str = "read") -> None:
    raw = np.asarray(bases, np.uint8).tobytes()
    with open(path, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">" + prefix.encode() + b"%d\n" % i + raw[int(off[i]):int(off[i + 1])] + b"\n")


def materialize_mhc(gold_dir: str, out_dir: str):
    """MHC_4 panel + CHM13 reads (BASELINE configs 1 and 2 with the bundled reads) -> (gfa path, reads path)."""
    z = np.load(os.path.join(gold_dir, "sketch_mhc4_chm13.npz"))
    k = np.load(os.path.join(gold_dir, "mhc4_panel_links.npz"))
    gfa = os.path.join(out_dir, "mhc4_panel.gfa")
    fa = os.path.join(out_dir, "chm13_reads.fa")
    write_gfa(gfa, z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], k["link_src"], k["link_dst"],
              [s.decode() if isinstance(s, bytes) else str(s) for s in k["walk_sample"].tolist()], k["walk_hap"])
    write_fasta(fa, z["read_bases"], z["read_off"])
    return gfa, fa


def materialize_mhc_replicated(gold_dir: str, out_dir: str, times: int):
    """The MHC_4 panel with every walk written `times` times (4 x = 20 W-lines, 18 x = 90 W-lines — identical
    lanes, the hardest case for the DP's tie-break under a different evaluation order) + the CHM13 reads
    -> (gfa path, reads path).  Copy r of a walk is sample "<name>_r<r>" with the original haplotype index."""
    z = np.load(os.path.join(gold_dir, "sketch_mhc4_chm13.npz"))
    k = np.load(os.path.join(gold_dir, "mhc4_panel_links.npz"))
    names = [s.decode() if isinstance(s, bytes) else str(s) for s in k["walk_sample"].tolist()]
    H = len(names)
    walk_vtx, walk_off = np.asarray(z["walk_vtx"]), np.asarray(z["walk_off"])
    vtx = np.concatenate([walk_vtx[int(walk_off[h]):int(walk_off[h + 1])] for _ in range(times) for h in range(H)])
    lens = np.diff(walk_off)
    off = np.concatenate([[0], np.cumsum(np.tile(lens, times))])
    gfa = os.path.join(out_dir, "mhc4_panel_x%d.gfa" % times)
    fa = os.path.join(out_dir, "chm13_reads.fa")
    write_gfa(gfa, z["seg_bases"], z["seg_off"], vtx, off, k["link_src"], k["link_dst"],
              [names[h] + ("_r%d" % r if r else "") for r in range(times) for h in range(H)], np.tile(np.asarray(k["walk_hap"]), times))
    write_fasta(fa, z["read_bases"], z["read_off"])
    return gfa, fa


def materialize_mhc_hg002_reads(gold_dir: str, out_dir: str, seed: int = 20261018, coverage: float = 2.0):
    """BASELINE config 2 with the documented substitute for the absent HG002 2x read set: 150-bp reads drawn
    uniformly from the HG002.1 and HG002.2 walk sequences of the MHC_4 panel (half each), random strand, 0.1 %
    substitutions, 2 x 5 Mbp in total (about 66 000 reads), seeded.  Returns (gfa path, reads path)."""
    from . import simulate
    z = np.load(os.path.join(gold_dir, "sketch_mhc4_chm13.npz"))
    k = np.load(os.path.join(gold_dir, "mhc4_panel_links.npz"))
    gfa, _ = materialize_mhc(gold_dir, out_dir)
    seg_bases, seg_off = np.asarray(z["seg_bases"], np.uint8), z["seg_off"]
    segs = [seg_bases[int(seg_off[v]):int(seg_off[v + 1])] for v in range(len(seg_off) - 1)]
    walk_vtx, walk_off = z["walk_vtx"], z["walk_off"]
    walks = [np.asarray(walk_vtx[int(walk_off[h]):int(walk_off[h + 1])]).tolist() for h in range(len(walk_off) - 1)]
    names = [s.decode() if isinstance(s, bytes) else str(s) for s in k["walk_sample"].tolist()]
    target = [h for h, n in enumerate(names) if n == "HG002"]
    if len(target) != 2:
        raise RuntimeError(f"expected the two HG002 walks in the MHC_4 panel, found {target}")
    reads = simulate.reads_from_walks(seed, dict(segs=segs, walks=walks, links=[]), target, coverage=coverage)
    fa = os.path.join(out_dir, "hg002_sim_reads.fa")
    simulate.write_reads(fa, reads)
    return gfa, fa


def materialize_vcf_panel(gold_dir: str, out_dir: str) -> str:
    """BASELINE config 3's graph: the GFA that dipgenie_b200/vcf2gfa.py derives from the reference's test/MHC_4.vcf.gz +
    MHC-CHM13.0.fa.gz, stored as tests/golden/mhc4_vcf_panel.npz (tests/golden/make_config3_golden.py) -> gfa path."""
    k = np.load(os.path.join(gold_dir, "mhc4_vcf_panel.npz"))
    gfa = os.path.join(out_dir, "mhc4_vcf_panel.gfa")
    write_gfa(gfa, k["seg_bases"], k["seg_off"], k["walk_vtx"], k["walk_off"], k["link_src"], k["link_dst"],
              [s.decode() if isinstance(s, bytes) else str(s) for s in k["walk_sample"].tolist()], k["walk_hap"])
    return gfa


# The reference's two toy inputs (test/test.gfa + read.fa: 8 segments, 5 walks, one 19-bp read;
# test/test2.gfa + read2.fa: 4 segments, 2 walks, one 87-bp read), restated as data.
TOY = {
    "test": dict(
        segs=["ATCG", "ATC", "AAA", "ATAC", "TTAC", "TGAC", "GCAT", "CATG"],
        links=[(0, 1), (0, 2), (1, 3), (2, 3), (3, 4), (3, 5), (3, 6), (4, 7), (5, 7), (6, 7)],
        walks=[[0, 1, 3, 6, 7], [0, 1, 3, 5, 7], [0, 2, 3, 6, 7], [0, 2, 3, 5, 7], [0, 2, 3, 4, 7]],
        names=[("test_hap_1", 0), ("test_hap_2", 1), ("test_hap_3", 2), ("test_hap_4", 3), ("test_hap_4", 4)],
        reads=["ATCGATCATACTTACCATG"]),
    "test2": dict(
        segs=["ACGTCATGCAGTCGTAACGTAGTCGTCACAGTCAGTCGTAGCTA", "A", "T", "GTAGCGTCAGTCAGTCAGTCGTAGCGTAACGTCGTAGTCAGT"],
        links=[(0, 1), (0, 2), (1, 3), (2, 3)],
        walks=[[0, 1, 3], [0, 2, 3]],
        names=[("test_hap_1", 0), ("test_hap_2", 1)],
        reads=["ACGTCATGCAGTCGTAACGTAGTCGTCACAGTCAGTCGTAGCTATGTAGCGTCAGTCAGTCAGTCGTAGCGTAACGTCGTAGTCAGT"]),
}


def materialize_toy(name: str, out_dir: str):
    t = TOY[name]
    seg_off = np.concatenate([[0], np.cumsum([len(s) for s in t["segs"]])])
    seg_bases = np.frombuffer("".join(t["segs"]).encode(), np.uint8)
    walk_off = np.concatenate([[0], np.cumsum([len(w) for w in t["walks"]])])
    walk_vtx = np.array([v for w in t["walks"] for v in w], np.int32)
    gfa = os.path.join(out_dir, name + ".gfa")
    fa = os.path.join(out_dir, name + "_reads.fa")
    write_gfa(gfa, seg_bases, seg_off, walk_vtx, walk_off, [a for a, _ in t["links"]], [b for _, b in t["links"]],
              [n for n, _ in t["names"]], [h for _, h in t["names"]])
    roff = np.concatenate([[0], np.cumsum([len(r) for r in t["reads"]])])
    write_fasta(fa, np.frombuffer("".join(t["reads"]).encode(), np.uint8), roff)
    return gfa, fa
