"""Diagnostics: joins an ncu source-page CSV (SASS view) of one kernel with nvdisasm line info and prints where the samples,
executed instructions and no-instruction stalls are, per source function (line ranges) and per source line.
Usage: python scripts/ncu_by_line.py <src.csv from `ncu -i rep --page source --csv`> <annotated sass from `nvdisasm -g -c`> <source file>"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, source = sys.argv[1:4]
addr_line = {}
cur = None
for ln in open(sass):
    m = re.search(r'line (\d+)', ln) if '//## File' in ln else None
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iex, inoi = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_no_inst")
ithr = hdr.index("Thread Instructions Executed")
base = None
per_line = defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows[2:]:
    if len(r) <= inoi: continue
    a = int(r[ia], 16)
    if base is None: base = a
    line = addr_line.get(a - base)
    v = per_line[line]
    v[0] += int(r[isamp] or 0); v[1] += int(r[iex] or 0); v[2] += int(r[inoi] or 0); v[3] += 1; v[4] += int(r[ithr] or 0)
# function ranges from the source: a line that starts a __device__/__global__ function
funcs = []
for i, ln in enumerate(open(source), 1):
    m = re.search(r'__(?:device|global)__.*?\b(\w+)\s*\(', ln)
    if m and not ln.strip().startswith('//'):
        funcs.append((i, m.group(1)))
def func_of(line):
    name = "?"
    for s, n in funcs:
        if line is not None and s <= line: name = n
        else: break
    return name
per_fn = defaultdict(lambda: [0, 0, 0, 0, 0])
for line, v in per_line.items():
    f = per_fn[func_of(line)]
    for k in range(5): f[k] += v[k]
tot = [sum(v[k] for v in per_line.values()) for k in range(5)]
print(f"total: samples {tot[0]} inst_executed {tot[1]} no_inst {tot[2]} sass {tot[3]} threads/inst {tot[4]/max(1,tot[1]):.1f}")
print(f"{'function':28s} {'samples%':>8s} {'exec%':>7s} {'no_inst%':>8s} {'sass':>6s} {'cyc/inst':>8s} {'thr/inst':>8s}")
for n, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:28s} {100*v[0]/tot[0]:8.1f} {100*v[1]/tot[1]:7.1f} {100*v[2]/max(1,tot[2]):8.1f} {v[3]:6d} {v[0]/max(1,v[1])*tot[1]/tot[0]:8.2f} {v[4]/max(1,v[1]):8.1f}")
print("top lines by samples:")
srcl = open(source).read().split('\n')
for line, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print(f"{str(line):>5s} {100*v[0]/tot[0]:5.1f}% exec {100*v[1]/tot[1]:5.1f}% noinst {100*v[2]/max(1,tot[2]):5.1f}% sass {v[3]:4d} | {srcl[line-1].strip()[:100] if line else ''}")
