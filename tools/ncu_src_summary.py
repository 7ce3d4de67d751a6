#!/usr/bin/env python
"""Summarise an ncu report's warp-state samples of field_fwd_kernel: per role (address range), top wait sites,
top instructions.  usage: ncu_src_summary.py report.ncu-rep"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = cur_line = None
by = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Line No", "Function Name"): continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == "" and len(r) > 4 and r[2].startswith("0x"):
        try: a = int(r[2], 16); c = int(r[4])
        except ValueError: continue
        e = by.setdefault(a, {"c": c, "t": r[3].strip(), "lines": []}); e["lines"].append((cur_file, cur_line))
addrs = sorted(by); base = addrs[0]
tot = sum(by[a]["c"] for a in addrs)
print("unique instructions", len(addrs), "samples", tot)
src = open("/root/repo/lab4d_b200/csrc/field_fwd.cu").read().split("\n")
def own_lines(a): return sorted({l for f, l in by[a]["lines"] if f == "field_fwd.cu" and l >= 96})
def ctx(i):
    ls = set()
    for j in range(max(0, i - 10), min(len(addrs), i + 10)): ls.update(own_lines(addrs[j]))
    return sorted(ls)
# wait sites
waits = [(by[a]["c"], i) for i, a in enumerate(addrs) if any(f == "ptx.cuh" for f, l in by[a]["lines"]) and ("SYNCS" in by[a]["t"] or "BRA" in by[a]["t"] or "NANOSLEEP" in by[a]["t"])]
wtot = sum(c for c, i in waits)
print(f"mbarrier wait samples {wtot} ({100*wtot/tot:.1f}%)")
waits.sort(reverse=True)
for c, i in waits[:16]:
    if c < 50: break
    print(f"  {c:6d} {100*c/tot:5.1f}% @{addrs[i]-base:#07x} ctx lines {ctx(i)[-8:]}")
# top non-wait instructions
wset = {addrs[i] for c, i in waits}
top = sorted(((by[a]["c"], a) for a in addrs if a not in wset), reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("top non-wait instructions")
for c, a in top:
    print(f"  {c:6d} {100*c/tot:5.1f}% @{a-base:#07x} {by[a]['t'][:58]:58s} {own_lines(a)[:4]}")
# per-role totals by address range (roles are laid out contiguously: producer, MMA issuers, compute)
def first_addr(pred):
    for a in addrs:
        if any(pred(l) for l in own_lines(a)): return a
    return None
marks = {}
for i, s in enumerate(src, 1):
    if "TMA producer" in s: marks["producer"] = i
    if "MMA issuers" in s: marks["mma"] = i
    if "compute / epilogue warps" in s: marks["compute"] = i
a_prod = first_addr(lambda l: marks["producer"] < l < marks["mma"])
a_mma = first_addr(lambda l: marks["mma"] < l < marks["compute"])
a_cmp = first_addr(lambda l: l > marks["compute"] + 3)
rng = collections.Counter(); rngw = collections.Counter()
for a in addrs:
    role = "init" if a < a_prod else "producer" if a < a_mma else "mma" if a < a_cmp else "compute"
    rng[role] += by[a]["c"]
    if a in wset: rngw[role] += by[a]["c"]
print("by role (samples, of which mbarrier waits):", {k: (v, rngw[k]) for k, v in rng.items()})
if len(sys.argv) > 3 and sys.argv[3] in ("producer", "mma", "compute"):
    role = sys.argv[3]
    lo, hi = {"producer": (a_prod, a_mma), "mma": (a_mma, a_cmp), "compute": (a_cmp, addrs[-1] + 1)}[role]
    print("instructions of role", role, "with >= 0.05% samples")
    for a in addrs:
        if lo <= a < hi and by[a]["c"] >= max(1, tot // 2000):
            print(f"  {by[a]['c']:6d} @{a-base:#07x} {by[a]['t'][:70]:70s} {own_lines(a)[:3]}")
if len(sys.argv) > 4:
    # histogram of compute-role samples by the *innermost* own source line (last in list), bucketed by 10 lines
    h = collections.Counter()
    for a in addrs:
        if a >= a_cmp:
            ls = own_lines(a)
            key = (ls[-1] // 10 * 10) if ls else -1
            h[key] += by[a]["c"]
    for k in sorted(h):
        if h[k] >= tot // 500: print(f"  lines {k:4d}-{k+9:4d}: {h[k]:6d}  {src[k].strip()[:90] if k > 0 else ''}")
