"""Regenerates tests/golden/mhc4_panel_links.npz (the L lines and walk names of test/MHC_4.gfa.gz as id arrays) and
checks, with the UNMODIFIED reference binary (oracle/_ref/DipGenie), that the GFA / FASTA files materialised from
the committed fixtures (dipgenie_b200/fixtures.py) give byte-identical output to the reference's own test files.
Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_e2e_inputs.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from dipgenie_b200 import fixtures  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "DipGenie")
REF_TEST = "/root/reference/test"


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def run_ref(gfa, reads, out, *flags):
    subprocess.run([REF_BIN, "-g", gfa, "-r", reads, "-o", out, "-t8", *flags], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    return md5(out)


def main():
    ids, src, dst, samples, haps = {}, [], [], [], []
    with gzip.open(os.path.join(REF_TEST, "MHC_4.gfa.gz"), "rt") as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[0] == "S":
                ids.setdefault(t[1], len(ids))
            elif t[0] == "L":
                a = ids.setdefault(t[1], len(ids))
                b = ids.setdefault(t[3], len(ids))
                assert t[2] == "+" and t[4] == "+"
                src.append(a); dst.append(b)
            elif t[0] == "W":
                samples.append(t[1]); haps.append(int(t[2]))
    z = np.load(os.path.join(HERE, "sketch_mhc4_chm13.npz"))
    assert len(ids) == len(z["seg_off"]) - 1
    np.savez_compressed(os.path.join(HERE, "mhc4_panel_links.npz"), link_src=np.array(src, np.int32), link_dst=np.array(dst, np.int32),
                        walk_sample=np.array(samples), walk_hap=np.array(haps, np.int32))
    exp = {}
    with tempfile.TemporaryDirectory() as td:
        gfa, fa = fixtures.materialize_mhc(HERE, td)
        for name, flags in (("mhc_p1", ["-p1"]), ("mhc_p2_R18", ["-p2", "-R18"]), ("mhc_p2_R6", ["-p2", "-R6"])):
            ours = run_ref(gfa, fa, os.path.join(td, "a.fa"), *flags)
            theirs = run_ref(os.path.join(REF_TEST, "MHC_4.gfa.gz"), os.path.join(REF_TEST, "CHM13_reads.fq.gz"), os.path.join(td, "b.fa"), *flags)
            assert ours == theirs, (name, ours, theirs)
            exp[name] = ours
        for toy, gf, rf, cases in (("test", "test.gfa", "read.fa", [["-p1", "-R2", "-k3", "-w2"], ["-p2", "-R2", "-k3", "-w2"], ["-p2", "-R2", "-k5", "-w3"],
                                                                      ["-p2", "-R0", "-k3", "-w2"]]),
                                   ("test2", "test2.gfa", "read2.fa", [["-p1", "-R2"], ["-p2", "-R2"]])):
            g2, f2 = fixtures.materialize_toy(toy, td)
            for flags in cases:
                ours = run_ref(g2, f2, os.path.join(td, "a.fa"), *flags)
                theirs = run_ref(os.path.join(REF_TEST, gf), os.path.join(REF_TEST, rf), os.path.join(td, "b.fa"), *flags)
                assert ours == theirs, (toy, flags, ours, theirs)
                exp[toy + " " + " ".join(flags)] = ours
    json.dump(exp, open(os.path.join(HERE, "e2e_expected.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(exp, indent=1))


if __name__ == "__main__":
    main()
