"""TEST INFRASTRUCTURE (runs the reference binary under oracle/_ref as the checker / baseline).  Wall time of the whole run: this repo's CLI (GPU) vs the unmodified reference binary (CPU), same inputs."""
import hashlib, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from dipgenie_b200 import _build, fixtures
td = tempfile.mkdtemp()
config2 = len(sys.argv) > 1 and sys.argv[1] == "--config2"      # HG002 read substitute (66 607 seeded reads) instead of the CHM13 reads
gfa, fa = fixtures.materialize_mhc_hg002_reads("tests/golden", td) if config2 else fixtures.materialize_mhc("tests/golden", td)
ncpu = os.cpu_count()
res = {}
for name, exe in (("dipgenie_b200", _build.CLI_BIN), ("reference", "oracle/_ref/DipGenie")):
    if not os.path.exists(exe):
        continue
    for flags in ((["-p2", "-R18"],) if config2 else (["-p2", "-R18"], ["-p1"])):
        out = os.path.join(td, name + ".fa")
        best = None
        for rep in range(2 if name == "dipgenie_b200" else 1):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", out, "-t%d" % ncpu, *flags], capture_output=True, text=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[name + " " + " ".join(flags)] = dict(seconds=round(best, 3), md5=hashlib.md5(open(out, "rb").read()).hexdigest(), rc=p.returncode)
        if name == "dipgenie_b200":
            res[name + " " + " ".join(flags)]["log"] = [l for l in p.stderr.splitlines() if l.startswith("[M::")][-6:]
print(json.dumps(dict(host_cpus=ncpu, reads="HG002 substitute (seed 20261018)" if config2 else "CHM13", runs=res), indent=1))
