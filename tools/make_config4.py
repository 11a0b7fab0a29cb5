"""BASELINE config 4: the synthetic H = 90 panel and its 30x read set, seeded (seed 90): backbone 5 Mbp,
33 000 biallelic sites (85 % SNP / 15 % indel), 12 founders, 90 mosaic walks, target diploid = two mosaics of panel
walks with 9 switches each, 150-bp reads at 30x (1 000 000 reads), 0.1 % substitutions, random strand.
Usage: make_config4.py OUT_DIR [scale]     (scale < 1 shrinks backbone, sites and reads alike)
Writes OUT_DIR/c4_h90.gfa and OUT_DIR/c4_h90_reads.fa."""
import os, sys, time
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import simulate


def make(out_dir, scale=1.0, walks=90, coverage=30.0, seed=90):
    os.makedirs(out_dir, exist_ok=True)
    t0 = time.time()
    backbone, sites = int(5_000_000 * scale), int(33_000 * scale)
    panel = simulate.make_panel(seed, backbone=backbone, n_sites=sites, n_founders=12, n_walks=walks, indel_frac=0.15,
                                breaks_per_walk=backbone / 250_000.0)
    gfa = os.path.join(out_dir, "c4_h%d.gfa" % walks)
    simulate.write_gfa(gfa, panel)
    t1 = time.time()
    reads = simulate.make_reads(seed, panel, coverage=coverage, switches=9)
    fa = os.path.join(out_dir, "c4_h%d_reads.fa" % walks)
    simulate.write_reads(fa, reads)
    print("config 4: %d segments, %d walks, %d reads; panel %.1f s, reads %.1f s" % (len(panel["segs"]), walks, len(reads), t1 - t0, time.time() - t1), file=sys.stderr)
    return gfa, fa


if __name__ == "__main__":
    make(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
