"""Seeded synthetic panels in the shape SURVEY 8d names for configs 4 and 5 (scaled by the caller): a random
backbone, biallelic SNP/indel sites, a few founder haplotypes, panel walks that are mosaics of founders, and
reads drawn from a diploid target that is itself a mosaic of panel walks.  Writes GFA 1.1 (S/L/W lines, forward
strand only, acyclic) and FASTA.  Test/bench input only — no reference code or data involved."""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def _rand_seq(rng, n):
    return _ACGT[rng.integers(0, 4, n)]


def make_panel(seed: int, backbone: int = 60000, n_sites: int = 400, n_founders: int = 6, n_walks: int = 12,
               indel_frac: float = 0.15, breaks_per_walk: float = 3.0):
    """Returns dict(segs, links, walks): segs = list of uint8 arrays; links = list of (u, v); walks = list of id lists."""
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(np.arange(5, backbone - 5, 3), size=n_sites, replace=False))   # >= 3 bp between sites
    base = _rand_seq(rng, backbone)
    segs, links = [], []
    piece_id, ref_id, alt_id = [], [], []
    prev_end = 0
    for s in range(n_sites):
        p = int(pos[s])
        is_indel = rng.random() < indel_frac
        ref_len = 1
        if is_indel and rng.random() < 0.5:
            ref_len = min(int(rng.geometric(1 / 12.0)) + 1, int((pos[s + 1] if s + 1 < n_sites else backbone - 5) - p - 1), 200)
            ref_len = max(ref_len, 1)
        piece_id.append(len(segs)); segs.append(base[prev_end:p])
        ref = base[p:p + ref_len]
        if is_indel and ref_len == 1:
            alt = np.concatenate([ref, _rand_seq(rng, min(int(rng.geometric(1 / 12.0)) + 1, 200))])     # insertion
        elif is_indel:
            alt = ref[:1]                                                                              # deletion
        else:
            alt = _ACGT[[(int(np.searchsorted(_ACGT, ref[0])) + 1 + int(rng.integers(0, 3))) % 4]]      # SNP
        ref_id.append(len(segs)); segs.append(ref)
        alt_id.append(len(segs)); segs.append(alt)
        prev_end = p + ref_len
    tail = len(segs); segs.append(base[prev_end:])
    for s in range(n_sites):
        nxt = piece_id[s + 1] if s + 1 < n_sites else tail
        links += [(piece_id[s], ref_id[s]), (piece_id[s], alt_id[s]), (ref_id[s], nxt), (alt_id[s], nxt)]
    af = rng.beta(0.3, 0.9, n_sites)
    founders = rng.random((n_founders, n_sites)) < af[None, :]
    founders[0, :] = False                                  # one founder is the backbone itself
    alleles = np.zeros((n_walks, n_sites), bool)
    for h in range(n_walks):
        nb = rng.poisson(breaks_per_walk)
        cuts = np.sort(rng.integers(0, n_sites, nb))
        f = int(rng.integers(0, n_founders))
        start = 0
        for c in list(cuts) + [n_sites]:
            alleles[h, start:c] = founders[f, start:c]
            f = int(rng.integers(0, n_founders)); start = c
    walks = []
    for h in range(n_walks):
        w = []
        for s in range(n_sites):
            w += [piece_id[s], alt_id[s] if alleles[h, s] else ref_id[s]]
        w.append(tail)
        walks.append(w)
    return dict(segs=segs, links=links, walks=walks, alleles=alleles)


def walk_sequence(panel, walk):
    return np.concatenate([panel["segs"][v] for v in walk])


def make_reads(seed: int, panel, coverage: float = 4.0, read_len: int = 150, err: float = 0.001, switches: int = 3):
    """Reads from a diploid target: each target haplotype is a mosaic of panel walks with `switches` switches."""
    rng = np.random.default_rng(seed + 7919)
    n_sites = panel["alleles"].shape[1]
    H = len(panel["walks"])
    reads = []
    for t in range(2):
        cuts = np.sort(rng.integers(0, n_sites, switches))
        al = np.zeros(n_sites, bool)
        start = 0
        for c in list(cuts) + [n_sites]:
            al[start:c] = panel["alleles"][int(rng.integers(0, H)), start:c]
            start = c
        w = panel["walks"][0]
        seq = []
        for s in range(n_sites):
            seq.append(panel["segs"][w[2 * s]])
            # ref/alt ids are consecutive after the piece (make_panel): piece, ref, alt
            seq.append(panel["segs"][w[2 * s] + (2 if al[s] else 1)])
        seq.append(panel["segs"][w[-1]])
        hap = np.concatenate(seq)
        n = int(coverage / 2 * len(hap) / read_len)
        starts = rng.integers(0, max(1, len(hap) - read_len), n)
        for st in starts:
            r = hap[st:st + read_len].copy()
            m = rng.random(len(r)) < err
            r[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
            if rng.random() < 0.5:
                r = _COMP[r[::-1]]
            reads.append(r)
    return reads


def reads_from_walks(seed: int, panel, walk_ids, coverage: float = 4.0, read_len: int = 150, err: float = 0.001):
    """Reads drawn evenly from the given panel walks (leave-one-out study: the two walks of the held-out sample)."""
    rng = np.random.default_rng(seed + 104729)
    reads = []
    for h in walk_ids:
        hap = walk_sequence(panel, panel["walks"][h])
        n = int(coverage / len(walk_ids) * len(hap) / read_len)
        for st in rng.integers(0, max(1, len(hap) - read_len), n):
            r = hap[st:st + read_len].copy()
            m = rng.random(len(r)) < err
            r[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
            if rng.random() < 0.5:
                r = _COMP[r[::-1]]
            reads.append(r)
    return reads


def without_walks(panel, drop):
    """The panel minus some walks (segments and links stay: the held-out sample's private alleles remain as
    vertices no walk visits, as in a graph built from the full cohort)."""
    keep = [h for h in range(len(panel["walks"])) if h not in set(drop)]
    return dict(segs=panel["segs"], links=panel["links"], walks=[panel["walks"][h] for h in keep],
                alleles=panel["alleles"][keep])


def write_gfa(path, panel, samples=None):
    with open(path, "wb") as f:
        f.write(b"H\tVN:Z:1.1\n")
        for v, s in enumerate(panel["segs"]):
            f.write(b"S\t%d\t" % (v + 1) + bytes(s) + b"\n")
        for a, b in panel["links"]:
            f.write(b"L\t%d\t+\t%d\t+\t0M\n" % (a + 1, b + 1))
        for h, w in enumerate(panel["walks"]):
            name = samples[h] if samples else "S%02d" % (h // 2 + 1)
            ln = sum(len(panel["segs"][v]) for v in w)
            f.write(b"W\t" + name.encode() + b"\t%d\tchr\t0\t%d\t" % (h % 2 + 1, ln) + b"".join(b">%d" % (v + 1) for v in w) + b"\n")


def write_reads(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i + bytes(r) + b"\n")
