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
