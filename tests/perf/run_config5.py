"""TEST INFRASTRUCTURE (runs the reference binary under oracle/_ref as the checker / baseline).  SURVEY 8d config 5 at reduced scale: 22-sample leave-one-out batch from a seeded 24-sample (48-walk) synthetic
panel.  Runs all 22 jobs in ONE `dipgenie -B` process (diploid DPs side by side on the GPU) and the unmodified
reference binary on the first `--ref-samples` jobs (one process per sample, all host threads, like
data/run_DipGenie_batch.sh); checks md5 parity on those and reports samples/s of both."""
import argparse, hashlib, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import _build, simulate

ap = argparse.ArgumentParser()
ap.add_argument("--backbone", type=int, default=60000)
ap.add_argument("--sites", type=int, default=500)
ap.add_argument("--samples", type=int, default=22)
ap.add_argument("--ref-samples", type=int, default=4)
ap.add_argument("--coverage", type=float, default=4.0)
ap.add_argument("--R", type=int, default=18)
a = ap.parse_args()
td = tempfile.mkdtemp()
panel = simulate.make_panel(22, backbone=a.backbone, n_sites=a.sites, n_walks=48, n_founders=12)
jobs = []
for s in range(a.samples):
    sub = simulate.without_walks(panel, [2 * s, 2 * s + 1])
    g, r, o = f"{td}/s{s}.gfa", f"{td}/s{s}.fa", f"{td}/s{s}.out.fa"
    simulate.write_gfa(g, sub)
    simulate.write_reads(r, simulate.reads_from_walks(500 + s, panel, [2 * s, 2 * s + 1], coverage=a.coverage))
    jobs.append((g, r, o))
open(f"{td}/jobs.tsv", "w").write("".join(f"{g}\t{r}\t{o}\n" for g, r, o in jobs))
ncpu = os.cpu_count()
t0 = time.perf_counter()
p = subprocess.run([_build.CLI_BIN, "-B", f"{td}/jobs.tsv", "-p2", f"-R{a.R}", f"-t{ncpu}"], capture_output=True, text=True)
t_ours = time.perf_counter() - t0
assert p.returncode == 0, p.stderr[-2000:]
batch_line = [l for l in p.stderr.splitlines() if "jobs in" in l]
t_ref, ok = 0.0, True
for g, r, o in jobs[: a.ref_samples]:
    t1 = time.perf_counter()
    subprocess.run(["oracle/_ref/DipGenie", "-g", g, "-r", r, "-o", o + ".ref", f"-t{ncpu}", "-p2", f"-R{a.R}"], capture_output=True, text=True, check=True)
    t_ref += time.perf_counter() - t1
    ok &= hashlib.md5(open(o, "rb").read()).hexdigest() == hashlib.md5(open(o + ".ref", "rb").read()).hexdigest()
print(json.dumps(dict(samples=a.samples, walks_per_graph=46, backbone=a.backbone, sites=a.sites, R=a.R, host_cpus=ncpu,
                      ours_wall_s=round(t_ours, 2), ours_samples_per_s=round(a.samples / t_ours, 2), ours_log=batch_line,
                      reference_s_per_sample=round(t_ref / max(1, a.ref_samples), 2),
                      reference_samples_per_s=round(a.ref_samples / t_ref, 4) if t_ref else None,
                      md5_equal_on_reference_samples=ok)))
