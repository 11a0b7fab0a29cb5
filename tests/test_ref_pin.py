"""Pins of the test infrastructure itself (build container only: needs /root/reference and oracle/_ref):
the INSTRUMENTED reference (oracle/_ref/ref_driver: the reference's objects plus a scratch copy of approximator.cpp
with read-only dump hooks, oracle/ref_hook.h) writes byte-identical FASTA to the UNMODIFIED reference binary
(oracle/_ref/DipGenie) — so the goldens dumped through the hooks are the reference's own results."""
import hashlib
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.ref
REF = os.path.join(ROOT, "oracle", "_ref")
TEST = "/root/reference/test"


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("gfa,reads,flags", [
    ("test.gfa", "read.fa", ["-p2", "-R2", "-k5", "-w3"]),
    ("test.gfa", "read.fa", ["-p1", "-R2", "-k3", "-w2"]),
    ("test2.gfa", "read2.fa", ["-p2", "-R2"]),
])
def test_instrumented_reference_writes_the_same_fasta(gfa, reads, flags, tmp_path):
    plain, instr, dump = tmp_path / "plain.fa", tmp_path / "instr.fa", tmp_path / "dump.dgd"
    common = ["-g", os.path.join(TEST, gfa), "-r", os.path.join(TEST, reads), "-t2"] + flags
    subprocess.run([os.path.join(REF, "DipGenie")] + common + ["-o", str(plain)], check=True, capture_output=True, timeout=600)
    subprocess.run([os.path.join(REF, "ref_driver")] + common + ["-o", str(instr), "-D", str(dump)], check=True, capture_output=True, timeout=600)
    assert os.path.getsize(plain) > 0 and md5(plain) == md5(instr)
    assert os.path.getsize(dump) > 0
