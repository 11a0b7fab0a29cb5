"""Small pure-Python readers used by tests and tools (the product front end is the C++ host library)."""
from __future__ import annotations

import gzip

import numpy as np


def _open(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def read_gfa_segments(path):
    """Segment sequences in the reference's id order (first mention on an S or L line,
    src/gfa-base.cpp:75-96).  Returns (names, seqs)."""
    ids = {}
    seqs = []

    def add(name):
        if name not in ids:
            ids[name] = len(seqs)
            seqs.append("")
        return ids[name]

    with _open(path) as f:
        for line in f:
            if len(line) < 3 or line[1] != "\t":
                continue
            t = line.rstrip("\n").split("\t")
            if t[0] == "S":
                i = add(t[1])
                seqs[i] = "" if t[2] == "*" else t[2]
            elif t[0] == "L":
                add(t[1])
                add(t[3])
    names = [None] * len(seqs)
    for n, i in ids.items():
        names[i] = n
    return names, seqs


def read_sequences(path):
    """FASTA/FASTQ(.gz) records -> list of sequences (kseq semantics: multi-line records allowed)."""
    out = []
    with _open(path) as f:
        lines = f.read().split("\n")
    i = 0
    n = len(lines)
    while i < n:
        ln = lines[i]
        if ln.startswith(">") or ln.startswith("@"):
            fastq = ln[0] == "@"
            i += 1
            seq = []
            while i < n and lines[i] and lines[i][0] not in ">+@":
                seq.append(lines[i])
                i += 1
            s = "".join(seq)
            out.append(s)
            if fastq and i < n and lines[i].startswith("+"):
                i += 1
                q = 0
                while i < n and q < len(s):
                    q += len(lines[i])
                    i += 1
        else:
            i += 1
    return out


def concat(seqs):
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer("".join(seqs).encode(), np.uint8) if len(seqs) else np.zeros(0, np.uint8)
    return bases, off
