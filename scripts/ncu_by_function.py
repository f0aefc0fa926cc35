"""Diagnostics: joins an ncu source-page CSV (SASS view) of one kernel with `nvdisasm -g -c` line info and prints where the
samples, executed instructions and no-instruction stalls are, per source function (several source files) and per line.
Usage: python scripts/ncu_by_function.py <src.csv from `ncu -i rep --page source --csv`> <nvdisasm -g -c output of the cubin>
       <mangled kernel name> [top lines] [substring of the kernel name in the report]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
addr_line = {}
cur = None
inside = False
for ln in open(sass):
    if ln.startswith("//--------------------- .text."):
        inside = ln.split(".text.")[1].split()[0] == kernel
        continue
    if not inside:
        continue
    if '//## File' in ln:
        m = re.search(r'File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
# a report with several kernels repeats {"Kernel Name", name} + header: keep the sections of the wanted kernel only
want = sys.argv[5] if len(sys.argv) > 5 else None
if want:
    keep, on = rows[:2], False
    for r in rows:
        if r and r[0] == "Kernel Name":
            on = want in r[1]
            continue
        if on and not (r and r[0] == "Address"):
            keep.append(r)
    rows = keep
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
inoi = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
ithr = hdr.index("Thread Instructions Executed")
base = None
per_line = defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows[2:]:
    if len(r) <= ithr or r[ia] == "Address":     # (a report with several launches repeats the header: the launches are summed)
        base = None if len(r) > ia and r[ia] == "Address" else base
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    v = per_line[addr_line.get(a - base)]
    v[0] += int(r[isamp] or 0); v[1] += int(r[iex] or 0); v[2] += int(r[inoi] or 0) if inoi is not None else 0; v[3] += 1
    v[4] += int(r[ithr] or 0)
sources = {}
funcs = {}
for f in {k[0] for k in per_line if k}:
    import glob
    path = glob.glob(f"tile_match_gym_b200/csrc/{f}")
    if not path:
        continue
    sources[f] = open(path[0]).read().split("\n")
    fl = []
    for i, ln in enumerate(sources[f], 1):
        m = re.search(r'__(?:device|global)__.*?\b(\w+)\s*\(', ln)
        if m and not ln.strip().startswith('//'):
            fl.append((i, m.group(1)))
    funcs[f] = fl


def func_of(key):
    if key is None or key[0] not in funcs:
        return "?"
    name = "?"
    for s, n in funcs[key[0]]:
        if s <= key[1]:
            name = n
        else:
            break
    return f"{key[0].replace('tmg_', '').replace('.cuh', '')}:{name}"


per_fn = defaultdict(lambda: [0, 0, 0, 0, 0])
for key, v in per_line.items():
    f = per_fn[func_of(key)]
    for k in range(5):
        f[k] += v[k]
tot = [sum(v[k] for v in per_line.values()) for k in range(5)]
print(f"total: samples {tot[0]} inst_executed {tot[1]} no_inst {tot[2]} sass {tot[3]} threads/inst {tot[4]/max(1,tot[1]):.1f}")
print(f"{'function':36s} {'samples%':>8s} {'exec%':>7s} {'no_inst%':>8s} {'sass':>6s} {'cyc/inst':>8s} {'thr/inst':>8s}")
for n, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:36s} {100*v[0]/tot[0]:8.1f} {100*v[1]/tot[1]:7.1f} {100*v[2]/max(1,tot[2]):8.1f} {v[3]:6d} {v[0]/max(1,v[1])*tot[1]/tot[0]:8.2f} {v[4]/max(1,v[1]):8.1f}")
print("top lines by samples:")
for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = sources[key[0]][key[1] - 1].strip()[:100] if key and key[0] in sources else ""
    print(f"{str(key):>24s} {100*v[0]/tot[0]:5.1f}% exec {100*v[1]/tot[1]:5.1f}% noinst {100*v[2]/max(1,tot[2]):5.1f}% sass {v[3]:4d} | {text}")
