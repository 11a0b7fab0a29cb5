"""VCF + reference FASTA -> GFA 1.1 with W lines (BASELINE config 3's input).

The reference's vcf2gfa.py (vcf2gfa.py:27-62) shells out to bgzip/tabix/samtools/vg/gfa2gbwt, none of which ship with it.
This is an own, deterministic converter with the same command line (`-v VCF -r FASTA`, GFA on stdout):

* the reference sequence is cut at the variant sites; every site becomes a bubble with one segment per allele that some
  haplotype carries (REF allele included: the reference walk carries it), backbone stretches between sites are one
  segment each;
* records that overlap an accepted record, and records with symbolic / `*` alleles or missing / unphased genotypes, are
  skipped (the haplotypes then follow the reference there);
* segments are numbered 1..n in creation (= topological) order; L lines connect everything that ends at a position
  with everything that starts there, in ascending id order; all orientations are '+', the graph is acyclic;
* walks: `W <REF sample> 0` for the reference sequence, then `W <sample> 1` / `W <sample> 2` per VCF sample from the
  phased GT column.

The output is a different (coarser) graph than `vg construct` would build, so parity for this config is reference binary
vs this repo's binary ON THIS GFA."""
from __future__ import annotations

import argparse
import gzip
import sys


def _open(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "rt")


def read_fasta(path):
    name, chunks = None, []
    with _open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    break                      # first sequence only
                name = line[1:].split()[0]
            else:
                chunks.append(line.strip().upper())
    return name, "".join(chunks)


def convert(vcf_path: str, ref_path: str, ref_sample: str = "CHM13"):
    """Returns dict(segs=[str], links=[(u, v)], walks=[[ids]], names=[(sample, hap)], skipped=int) with 0-based ids."""
    _, ref = read_fasta(ref_path)
    segs, links, samples = [], [], []
    walks = None
    frontier = []            # segments that end at `pos`
    pos = 0                  # next reference base (0-based) not yet emitted
    skipped = 0

    def add_seg(seq):
        segs.append(seq)
        return len(segs) - 1

    def connect(new_nodes):
        for a in frontier:
            for b in new_nodes:
                links.append((a, b))

    with _open(vcf_path) as f:
        for line in f:
            if line.startswith("##"):
                continue
            t = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                samples = t[9:]
                walks = [[] for _ in range(1 + 2 * len(samples))]
                continue
            start = int(t[1]) - 1
            alleles = [t[3].upper()] + t[4].upper().split(",")
            gts = [g.split(":")[0] for g in t[9:]]
            ok = (start >= pos and ref[start:start + len(alleles[0])] == alleles[0]
                  and all(a and set(a) <= set("ACGTN") for a in alleles)
                  and all(len(g.split("|")) == 2 and all(x.isdigit() and int(x) < len(alleles) for x in g.split("|")) for g in gts))
            if start == 0:
                ok = False                      # a site at the very first base would leave no source segment
            if not ok:
                skipped += 1
                continue
            hap_alleles = [0] + [int(x) for g in gts for x in g.split("|")]
            if start > pos:
                b = add_seg(ref[pos:start])
                connect([b])
                frontier = [b]
                for w in walks:
                    w.append(b)
            used = sorted(set(hap_alleles))
            node_of = {a: add_seg(alleles[a]) for a in used}
            connect([node_of[a] for a in used])
            frontier = [node_of[a] for a in used]
            for w, a in zip(walks, hap_alleles):
                w.append(node_of[a])
            pos = start + len(alleles[0])
    if pos < len(ref):
        b = add_seg(ref[pos:])
        connect([b])
        for w in walks:
            w.append(b)
    names = [(ref_sample, 0)] + [(s, h) for s in samples for h in (1, 2)]
    return dict(segs=segs, links=links, walks=walks, names=names, skipped=skipped)


def write_gfa(g, out):
    out.write("H\tVN:Z:1.1\n")
    for v, s in enumerate(g["segs"]):
        out.write(f"S\t{v + 1}\t{s}\n")
    for a, b in g["links"]:
        out.write(f"L\t{a + 1}\t+\t{b + 1}\t+\t0M\n")
    for (sample, hap), w in zip(g["names"], g["walks"]):
        ln = sum(len(g["segs"][v]) for v in w)
        out.write(f"W\t{sample}\t{hap}\tchr\t0\t{ln}\t" + "".join(f">{v + 1}" for v in w) + "\n")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Generate GFA from VCF and FASTA/FA files.")
    ap.add_argument("-v", "--vcf", required=True, help="Input VCF file (can be gzipped).")
    ap.add_argument("-r", "--ref", required=True, help="Input reference FASTA/FA file (can be gzipped).")
    ap.add_argument("--ref-sample", default="CHM13", help="sample name of the reference walk")
    a = ap.parse_args(argv)
    write_gfa(convert(a.vcf, a.ref, a.ref_sample), sys.stdout)


if __name__ == "__main__":
    main()
