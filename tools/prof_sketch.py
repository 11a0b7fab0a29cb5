"""Times the sketch stages on the bundled MHC input (diagnostics; prints JSON lines)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipgenie_b200.cuda_api import Context
z = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sketch_mhc4_chm13.npz")))   # materialise once: NpzFile decompresses on every access
k, w = int(z["k"]), int(z["w"])
with Context(0) as ctx:
    for rep in range(3):
        t0 = time.perf_counter(); sp, rc = ctx.sketch_reads(z["read_bases"], z["read_off"], k, w); t1 = time.perf_counter()
        s1 = ctx.sketch_stats()
        ix = ctx.index_walks(z["seg_bases"], z["seg_off"], z["walk_vtx"], z["walk_off"], z["top_order_map"], k, w, sp); t2 = time.perf_counter()
        s2 = ctx.sketch_stats()
    print(json.dumps({"reads": s1, "reads_wall_ms": (t1 - t0) * 1e3, "walks": s2, "walks_wall_ms": (t2 - t1) * 1e3,
                      "hits": [int(x) for x in np.diff(ix["hit_off"].astype(np.int64))]}))
