"""Seeded synthetic panels (dipgenie_b200/simulate.py: mosaic walks over SNP/indel sites, reads from a diploid
mosaic target — the shape of SURVEY 8d's configs 4/5, scaled down) run through the UNMODIFIED reference binary
(oracle/_ref/DipGenie, CPU) and through this repo: FASTA bytes must be identical.
  * `-m "not gpu"`: host glue + oracle stages (tests/host/host_check);
  * `-m gpu`: the dipgenie CLI with every hot-path stage on the GPU, including panels wide enough for the
    row-split transitions and for destinations with more than 32 in-edges (pair-form fallback)."""
import hashlib
import os
import subprocess

import pytest

from conftest import ROOT
from dipgenie_b200 import _build, simulate

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "DipGenie")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/DipGenie not built (make -C oracle ref)")

PANELS = {
    "h8": (dict(backbone=30000, n_sites=200, n_walks=8), 4.0),
    "h24": (dict(backbone=60000, n_sites=500, n_walks=24, n_founders=8), 6.0),
    "h40": (dict(backbone=40000, n_sites=300, n_walks=40, n_founders=10, breaks_per_walk=2.0), 8.0),
}


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def materialize(name, seed, tmp):
    kw, cov = PANELS[name]
    p = simulate.make_panel(seed, **kw)
    g, r = os.path.join(tmp, f"{name}.gfa"), os.path.join(tmp, f"{name}.fa")
    simulate.write_gfa(g, p)
    simulate.write_reads(r, simulate.make_reads(seed, p, coverage=cov))
    return g, r


def run_ref(g, r, out, flags):
    p = subprocess.run([REF_BIN, "-g", g, "-r", r, "-o", out, "-t8", *flags], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-500:]
    return md5(out)


@needs_ref
@pytest.mark.parametrize("name,seed,flags", [("h8", 1, ["-p2", "-R6"]), ("h8", 1, ["-p1"]), ("h24", 2, ["-p2", "-R6"])])
def test_host_glue_matches_reference_on_synthetic(name, seed, flags, tmp_path):
    import oracle
    oracle.build()
    _build.build_host()
    exe = os.path.join(ROOT, "tests", "host", "host_check")
    if not os.path.exists(exe):
        pytest.skip("tests/host/host_check not built (tests/test_host_cpu.py builds it)")
    g, r = materialize(name, seed, str(tmp_path))
    want = run_ref(g, r, str(tmp_path / "ref.fa"), flags)
    args = [exe, "-g", g, "-r", r, "-o", str(tmp_path / "ours.fa"), "-t", "8"]
    for f in flags:
        args += [f[:2], f[2:]]
    p = subprocess.run(args, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-500:]
    assert md5(str(tmp_path / "ours.fa")) == want


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,flags", [("h8", 3, ["-p2", "-R4"]), ("h8", 3, ["-p1"]), ("h24", 4, ["-p2", "-R18"]),
                                              ("h24", 4, ["-p1", "-R6"]), ("h40", 5, ["-p2", "-R6"]), ("h40", 6, ["-p2", "-R2"])])
def test_cli_matches_reference_on_synthetic(name, seed, flags, tmp_path):
    assert os.path.exists(_build.CLI_BIN)
    g, r = materialize(name, seed, str(tmp_path))
    want = run_ref(g, r, str(tmp_path / "ref.fa"), flags)
    p = subprocess.run([_build.CLI_BIN, "-g", g, "-r", r, "-o", str(tmp_path / "ours.fa"), "-t8", *flags], capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-1500:]
    assert md5(str(tmp_path / "ours.fa")) == want


@needs_ref
@pytest.mark.gpu
def test_cli_batch_leave_one_out_matches_reference(tmp_path):
    """SURVEY 8d config 5 at reduced scale: a 7-sample (14-walk) panel, every sample held out in turn (graph = panel
    minus its two walks, reads from its two walks); one `dipgenie -B` process runs all jobs with the diploid DPs side
    by side on the GPU; every FASTA must equal the reference binary's run on the same job."""
    assert os.path.exists(_build.CLI_BIN)
    panel = simulate.make_panel(11, backbone=30000, n_sites=220, n_walks=14, n_founders=6)
    jobs = []
    for s in range(7):
        sub = simulate.without_walks(panel, [2 * s, 2 * s + 1])
        g, r, o = str(tmp_path / f"loo{s}.gfa"), str(tmp_path / f"loo{s}.fa"), str(tmp_path / f"loo{s}.out.fa")
        simulate.write_gfa(g, sub)
        simulate.write_reads(r, simulate.reads_from_walks(100 + s, panel, [2 * s, 2 * s + 1], coverage=4.0))
        jobs.append((g, r, o))
    manifest = tmp_path / "jobs.tsv"
    manifest.write_text("".join(f"{g}\t{r}\t{o}\n" for g, r, o in jobs))
    p = subprocess.run([_build.CLI_BIN, "-B", str(manifest), "-p2", "-R6", "-t8"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-1500:]
    assert "7 jobs" in p.stderr
    for g, r, o in jobs:
        assert md5(o) == run_ref(g, r, o + ".ref", ["-p2", "-R6"])
