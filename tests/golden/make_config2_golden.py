"""Adds BASELINE config 2 with the documented HG002 read substitute (fixtures.materialize_mhc_hg002_reads: seeded reads
from the HG002.1 / HG002.2 walks of MHC_4) to tests/golden/e2e_expected.json: FASTA md5 of the UNMODIFIED reference
binary (oracle/_ref/DipGenie, `make -C oracle ref`) with -p2 -R18, at -t1 and -t8 (must agree).  Build container only:
    python tests/golden/make_config2_golden.py
"""
import hashlib, json, os, subprocess, sys, tempfile, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from dipgenie_b200 import fixtures  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "DipGenie")
exp_path = os.path.join(HERE, "e2e_expected.json")
exp = json.load(open(exp_path))
with tempfile.TemporaryDirectory() as td:
    gfa, fa = fixtures.materialize_mhc_hg002_reads(HERE, td)
    n_reads = sum(1 for line in open(fa, "rb") if line.startswith(b">"))
    md5s, log = {}, ""
    for t in ("-t8", "-t1"):
        out = os.path.join(td, "ref" + t + ".fa")
        t0 = time.perf_counter()
        p = subprocess.run([REF_BIN, "-g", gfa, "-r", fa, "-o", out, t, "-p2", "-R18"], capture_output=True, text=True, check=True)
        print(t, round(time.perf_counter() - t0, 1), "s", file=sys.stderr)
        md5s[t] = hashlib.md5(open(out, "rb").read()).hexdigest()
        log = p.stdout + p.stderr
    assert md5s["-t8"] == md5s["-t1"], md5s
    exp["mhc_hg002sim_p2_R18"] = md5s["-t8"]
    exp["mhc_hg002sim_reads_md5"] = hashlib.md5(open(fa, "rb").read()).hexdigest()
    exp["mhc_hg002sim_n_reads"] = n_reads
    print([l for l in log.splitlines() if "DP value" in l or "ecombination" in l][:4], file=sys.stderr)
json.dump(exp, open(exp_path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: v for k, v in exp.items() if "hg002" in k}, indent=1))
