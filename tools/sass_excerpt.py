"""Counts of the SASS mnemonics that matter (TMA bulk copies, mbarrier ops, fused add-max, warp reductions, bulk L2 prefetch)
per kernel of libdipgenie_cuda.so -> markdown on stdout.  Usage: sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import collections, re, subprocess, sys
so = "dipgenie_b200/libdipgenie_cuda.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UBLKCP", "UBLKPF", "SYNCS", "VIADDMNMX", "REDUX", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "ATOM", "SHFL", "VOTE", "POPC", "UTMALDG", "UTCMMA", "HMMA"]
cur, tab, size = None, collections.OrderedDict(), {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); tab[cur] = collections.Counter(); size[cur] = 0; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        size[cur] += 1
        op = m.group(1)
        for w in want:
            if op == w or op.startswith(w + ".") or (w in ("BAR.SYNC", "ATOM") and op.startswith(w)) or (w == "REDUX" and "REDUX" in op):
                tab[cur][w] += 1
print("# SASS excerpt of `libdipgenie_cuda.so` (sm_100a) — mnemonic counts per kernel\n")
print("`cuobjdump -sass dipgenie_b200/libdipgenie_cuda.so`, counted by `tools/sass_excerpt.py`.  `UBLKCP` = `cp.async.bulk` (TMA 1-D bulk copy), `UBLKPF` = `cp.async.bulk.prefetch.L2`,")
print("`SYNCS` = mbarrier arrive / try_wait / expect_tx, `VIADDMNMX` = fused add + max of the packed-key candidates, `REDUX` = warp-wide integer maximum.")
print("No `UTMALDG` / `UTCMMA` / `HMMA`: the path is integer max-plus over contiguous records, there is no contraction to put on the tensor cores.\n")
cols = [w for w in want if any(tab[k][w] for k in tab)]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
def short(n):
    s = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    s = re.sub(r"\(.*$", "", s).replace("dg::", "").replace("void ", "")
    return s
for k in tab:
    if size[k] < 50:
        continue
    n = short(k)
    if n.startswith("cub::") or "thrust" in n:
        continue
    print(f"| `{n}` | {size[k]} | " + " | ".join(str(tab[k][w]) for w in cols) + " |")
