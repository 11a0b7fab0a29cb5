"""Dumps a level-graph npz to raw arrays, builds tools/plan_time.cpp and runs it (host only; no GPU)."""
import sys, os, subprocess, numpy as np
sys.path.insert(0, os.getcwd())
from dipgenie_b200.cuda_api import LevelGraph
npz = sys.argv[1] if len(sys.argv) > 1 else 'tests/golden/mhc4_chm13_dipin.npz'
out = '/tmp/plan_time_in'; os.makedirs(out, exist_ok=True)
z, _ = LevelGraph.from_npz(npz)
for name, dt in (("level_off", np.int32), ("adj_off", np.int64), ("adj_dst", np.int32), ("adj_w", np.uint8),
                 ("col_off", np.int64), ("col_val", np.int32), ("colour_is_hom", np.uint8)):
    np.ascontiguousarray(getattr(z, name), dtype=dt).tofile(f"{out}/{name}.bin")
src = 'dipgenie_b200/csrc/cuda'
subprocess.check_call(["g++", "-O3", "-fopenmp", "-std=c++17", "-I", src, "tools/plan_time.cpp", f"{src}/dp_prep.cpp", f"{src}/dp_plan4.cpp", "-o", "/tmp/plan_time"])
subprocess.check_call(["/tmp/plan_time", out] + sys.argv[2:])
