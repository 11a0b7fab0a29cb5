"""Times the haploid DP on the bundled MHC input (diagnostics; prints one JSON line)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipgenie_b200.cuda_api import Context, HapGraph
g = HapGraph.from_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mhc4_chm13_hapin.npz"))
with Context(0) as ctx:
    p = ctx.hap_create(g, 18)
    for _ in range(3):
        p.run(); r = p.result()
    print(json.dumps({"stats": p.stats(), "colours": r["colours_by_r"].tolist()[:3]}))
    p.close()
