"""BASELINE config 3 (diploid -p2 -R18 on a VCF-derived graph): runs this repo's own VCF -> GFA converter
(dipgenie_b200/vcf2gfa.py) on the reference's test/MHC_4.vcf.gz + test/MHC-CHM13.0.fa.gz, stores the graph as the compact
fixture tests/golden/mhc4_vcf_panel.npz (so that it travels to the GPU box), and records the FASTA md5 of the UNMODIFIED
reference binary on the GFA materialised from that fixture with the config-2 read substitute.  Build container only:
    python tests/golden/make_config3_golden.py
"""
import hashlib, json, os, subprocess, sys, tempfile, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from dipgenie_b200 import fixtures, vcf2gfa  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "DipGenie")
REF_TEST = "/root/reference/test"
g = vcf2gfa.convert(os.path.join(REF_TEST, "MHC_4.vcf.gz"), os.path.join(REF_TEST, "MHC-CHM13.0.fa.gz"))
seg_off = np.concatenate([[0], np.cumsum([len(s) for s in g["segs"]])]).astype(np.int64)
np.savez_compressed(os.path.join(HERE, "mhc4_vcf_panel.npz"),
                    seg_bases=np.frombuffer("".join(g["segs"]).encode(), np.uint8), seg_off=seg_off,
                    walk_vtx=np.array([v for w in g["walks"] for v in w], np.int32),
                    walk_off=np.concatenate([[0], np.cumsum([len(w) for w in g["walks"]])]).astype(np.int64),
                    link_src=np.array([a for a, _ in g["links"]], np.int32), link_dst=np.array([b for _, b in g["links"]], np.int32),
                    walk_sample=np.array([n for n, _ in g["names"]]), walk_hap=np.array([h for _, h in g["names"]], np.int32))
exp_path = os.path.join(HERE, "e2e_expected.json")
exp = json.load(open(exp_path))
with tempfile.TemporaryDirectory() as td:
    gfa = fixtures.materialize_vcf_panel(HERE, td)
    _, fa = fixtures.materialize_mhc_hg002_reads(HERE, td)
    out = os.path.join(td, "ref.fa")
    t0 = time.perf_counter()
    p = subprocess.run([REF_BIN, "-g", gfa, "-r", fa, "-o", out, "-t8", "-p2", "-R18"], capture_output=True, text=True, check=True)
    print(round(time.perf_counter() - t0, 1), "s", file=sys.stderr)
    log = p.stdout + p.stderr
    print([l[-120:] for l in log.splitlines() if "DP value" in l or "ecombinations in" in l][:4], file=sys.stderr)
    exp["mhc_vcf_hg002sim_p2_R18"] = hashlib.md5(open(out, "rb").read()).hexdigest()
    exp["mhc_vcf_gfa_md5"] = hashlib.md5(open(gfa, "rb").read()).hexdigest()
    exp["mhc_vcf_segments"] = len(g["segs"])
    exp["mhc_vcf_skipped_records"] = g["skipped"]
json.dump(exp, open(exp_path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: v for k, v in exp.items() if "vcf" in k}, indent=1))
