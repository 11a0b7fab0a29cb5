"""Host glue end to end on a CPU-only box: the product's libdipgenie_host.so (GFA / read parsing, panel model,
anchor filter + hom/het classifier, graph expansion, Kahn order, levelization, stitching, FASTA) driven through
the same stage table as the CLI, with the oracle standing in for the four device stages (tests/host/host_check.cpp,
test infrastructure).  The FASTA files must be byte-identical to the reference binary's (md5s recorded by
tests/golden/make_e2e_inputs.py, which also proves that the materialised inputs equal the reference's test files)."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import GOLD, ROOT
from dipgenie_b200 import _build, fixtures


@pytest.fixture(scope="module")
def host_check():
    import oracle
    oracle.build()
    _build.build_host()
    exe = os.path.join(ROOT, "tests", "host", "host_check")
    src = os.path.join(ROOT, "tests", "host", "host_check.cpp")
    deps = [src, _build.HOST_LIB, os.path.join(ROOT, "oracle", "liboracle.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, src, "-L" + _build.PKG, "-ldipgenie_host",
                               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + _build.PKG,
                               "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lz", "-lm"])
    return exe


@pytest.fixture(scope="module")
def e2e_expected():
    return json.load(open(os.path.join(GOLD, "e2e_expected.json")))


def run(exe, gfa, reads, out, flags):
    args = [exe, "-g", gfa, "-r", reads, "-o", out, "-t", "8"]
    for f in flags:
        args += [f[:2], f[2:]]
    p = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    return json.loads(p.stdout.strip().splitlines()[-1]), hashlib.md5(open(out, "rb").read()).hexdigest()


@pytest.mark.parametrize("toy,flags", [("test", ["-p1", "-R2", "-k3", "-w2"]), ("test", ["-p2", "-R2", "-k3", "-w2"]),
                                        ("test", ["-p2", "-R2", "-k5", "-w3"]), ("test", ["-p2", "-R0", "-k3", "-w2"]),
                                        ("test2", ["-p1", "-R2"]), ("test2", ["-p2", "-R2"])])
def test_toy_inputs_match_reference(toy, flags, host_check, e2e_expected, tmp_path):
    gfa, fa = fixtures.materialize_toy(toy, str(tmp_path))
    s, md5 = run(host_check, gfa, fa, str(tmp_path / "out.fa"), flags)
    assert md5 == e2e_expected[toy + " " + " ".join(flags)]
    if toy == "test" and flags == ["-p2", "-R2", "-k5", "-w3"]:
        assert (s["dp_value"], s["r1"], s["r2"]) == (14, 1, 1)
        assert open(tmp_path / "out.fa").read().split("\n")[1::2][:2] == ["ATCGAAAATACTTACCATG", "ATCGATCATACGCATCATG"]


def test_mhc_diploid_matches_reference(host_check, e2e_expected, tmp_path):
    """BASELINE config 2's run (MHC_4 panel, bundled CHM13 reads, -p2 -R18): md5 46394489…, DP value 60729, 17/1."""
    gfa, fa = fixtures.materialize_mhc(GOLD, str(tmp_path))
    s, md5 = run(host_check, gfa, fa, str(tmp_path / "out.fa"), ["-p2", "-R18"])
    assert md5 == e2e_expected["mhc_p2_R18"] == "46394489af8bc9026605ddf237aca4c7"
    assert (s["dp_value"], s["r1"], s["r2"], s["len1"], s["len2"]) == (60729, 17, 1, 5005629, 4920284)
    assert s["spectrum"] == 138834 and round(100 * s["n_hom"] / (s["n_hom"] + s["n_het"]), 2) == 15.38


def test_mhc_haploid_matches_reference(host_check, e2e_expected, tmp_path):
    """BASELINE config 1 (-p1): md5 0c4df87d…, LN:4916718, recombination count 0."""
    gfa, fa = fixtures.materialize_mhc(GOLD, str(tmp_path))
    s, md5 = run(host_check, gfa, fa, str(tmp_path / "out.fa"), ["-p1"])
    assert md5 == e2e_expected["mhc_p1"] == "0c4df87ded10634a36db0a2c90521ff0"
    assert s["len1"] == 4916718 and s["best_r"] == 0


def test_config2_read_substitute_is_reproducible(tmp_path):
    """The seeded HG002 read substitute of BASELINE config 2 (fixtures.materialize_mhc_hg002_reads) must come out
    byte-identical wherever it is generated: its reference golden was recorded in the build container."""
    import hashlib
    import json
    from dipgenie_b200 import fixtures
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    exp = json.load(open(os.path.join(gold, "e2e_expected.json")))
    _, fa = fixtures.materialize_mhc_hg002_reads(gold, str(tmp_path))
    assert hashlib.md5(open(fa, "rb").read()).hexdigest() == exp["mhc_hg002sim_reads_md5"]
    assert sum(1 for line in open(fa, "rb") if line.startswith(b">")) == exp["mhc_hg002sim_n_reads"] == 66607


def test_vcf2gfa_small_cases(tmp_path):
    """Own VCF -> GFA converter (dipgenie_b200/vcf2gfa.py, replaces the external-tool pipeline of the reference's
    vcf2gfa.py): every walk must spell its haplotype; overlapping and symbolic records are skipped."""
    from dipgenie_b200 import vcf2gfa
    ref = "ACGTACGTTTGACCAGTAGGCATCGATTACA"
    (tmp_path / "r.fa").write_text(">chrT\n" + ref[:16] + "\n" + ref[16:] + "\n")
    recs = [(3, "G", "T", "0|1", "1|1"),            # SNP
            (6, "CGTT", "C", "1|0", "0|0"),         # deletion (anchor base kept)
            (8, "T", "TAAA", "0|1", "0|0"),         # overlaps the deletion: skipped
            (12, "A", "AGG,C", "1|2", "0|2"),       # multi-allelic: insertion and SNP
            (13, "C", "G", "1|1", "1|0"),           # adjacent site (no backbone between)
            (20, "G", "<DEL>", "0|1", "0|0"),       # symbolic: skipped
            (26, "A", "T", "0|0", "0|0")]           # nobody carries ALT: only the REF allele becomes a segment
    vcf = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
    for pos, r, a, g1, g2 in recs:
        assert ref[pos - 1:pos - 1 + len(r)] == r
        vcf += f"chrT\t{pos}\t.\t{r}\t{a}\t60\t.\t.\tGT\t{g1}\t{g2}\n"
    (tmp_path / "v.vcf").write_text(vcf)
    g = vcf2gfa.convert(str(tmp_path / "v.vcf"), str(tmp_path / "r.fa"), "REF")
    assert g["skipped"] == 2
    assert g["names"] == [("REF", 0), ("S1", 1), ("S1", 2), ("S2", 1), ("S2", 2)]

    def hap(col, idx):            # apply the accepted records to the reference, right to left
        s = ref
        for pos, r, a, g1, g2 in reversed([x for x in recs if x[0] not in (8, 20)]):
            al = int((g1, g2)[col].split("|")[idx])
            if al:
                s = s[:pos - 1] + a.split(",")[al - 1] + s[pos - 1 + len(r):]
        return s
    spelled = ["".join(g["segs"][v] for v in w) for w in g["walks"]]
    assert spelled[0] == ref
    assert spelled[1:] == [hap(0, 0), hap(0, 1), hap(1, 0), hap(1, 1)]
    for w in g["walks"]:          # every step of a walk is an L line; ids ascend (topological order)
        assert all((a, b) in set(g["links"]) and a < b for a, b in zip(w, w[1:]))
    assert sum(1 for s in g["segs"] if s == "T") >= 1 and len(g["segs"]) == len(set(v for w in g["walks"] for v in w))


def test_vcf2gfa_reproduces_the_config3_fixture(tmp_path):
    """tests/golden/mhc4_vcf_panel.npz (what the GPU box runs config 3 on) is exactly what the converter derives from the
    reference's test/MHC_4.vcf.gz + MHC-CHM13.0.fa.gz (only checkable where the reference tree exists)."""
    ref_test = "/root/reference/test"
    if not os.path.exists(os.path.join(ref_test, "MHC_4.vcf.gz")):
        pytest.skip("reference test data not present")
    from dipgenie_b200 import vcf2gfa
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    exp = json.load(open(os.path.join(gold, "e2e_expected.json")))
    gfa = fixtures.materialize_vcf_panel(gold, str(tmp_path))
    assert hashlib.md5(open(gfa, "rb").read()).hexdigest() == exp["mhc_vcf_gfa_md5"]
    g = vcf2gfa.convert(os.path.join(ref_test, "MHC_4.vcf.gz"), os.path.join(ref_test, "MHC-CHM13.0.fa.gz"))
    assert len(g["segs"]) == exp["mhc_vcf_segments"] and g["skipped"] == exp["mhc_vcf_skipped_records"]
    with open(tmp_path / "direct.gfa", "w") as f:
        vcf2gfa.write_gfa(g, f)
    strip = lambda p: [l for l in open(p) if l[0] in "SLW"]          # noqa: E731
    assert strip(tmp_path / "direct.gfa") == strip(gfa)
