"""Diagnostics: static SASS instruction count of a kernel per source function (nvdisasm -g -c line info), and the sizes of
all kernels / device functions of the cubin.  usage: python scripts/sass_by_function.py <nvdisasm -g -c output> <mangled kernel>"""
import re
import subprocess
import sys
from collections import defaultdict

sass, kernel = sys.argv[1:3]
srcs = {f: open('tile_match_gym_b200/csrc/' + f).read().split('\n') for f in ('tmg_device.cuh', 'tmg_rb.cuh')}
funcs = {}
for f, lines in srcs.items():
    fl = []
    for i, ln in enumerate(lines, 1):
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
    return key[0].replace('tmg_', '').replace('.cuh', '') + ':' + name


cnt = defaultdict(int)
sections = defaultdict(int)
cur = sec = None
sub = None
subs = defaultdict(int)
for ln in open(sass):
    if ln.startswith("//--------------------- .text."):
        sec = ln.split(".text.")[1].split()[0]
        continue
    m = re.match(r'\s*\.type\s+(\S+),@function', ln)
    if m:
        sub = m.group(1)
    if '//## File' in ln:
        m = re.search(r'File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        sections[sec] += 1
        if sec == kernel:
            subs[sub] += 1
            if sub == kernel:
                cnt[func_of(cur)] += 1
print("functions inside the kernel's section:")
for k, v in sorted(subs.items(), key=lambda kv: -kv[1]):
    print(f"{v:6d} {subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip().replace('tmg::', '')[:120]}")
print("kernel body by source function:")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v:6d} {k}")
