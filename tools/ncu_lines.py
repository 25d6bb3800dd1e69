"""Join an `ncu --page source --csv` SASS dump of one kernel with `nvdisasm -g -c` line info of the cubin it came
from, and print the hottest source lines (warp instructions executed, samples, mean active threads).

usage: python tools/ncu_lines.py <ncu_sass.csv> <nvdisasm.txt> <kernel substring> [top_n]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# nvdisasm: offset -> innermost "file:line" chain
line_of = {}
cur, active = None, False
for ln in open(dis):
    if ln.startswith('.text.') and ln.rstrip().endswith(':'):
        active = kname in ln
        continue
    if ln.startswith('//-----') and '.text.' in ln:
        active = kname in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = f'{m.group(1).split("/")[-1]}:{m.group(2)}'
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(sass_csv)))
hdr = next(r for r in rows if r and r[0] == 'Address')
body = []
for r in rows[rows.index(hdr) + 1:]:
    if not r or r[0] in ('Kernel Name', 'Address'):
        break
    body.append(r)
ia, isrc = hdr.index('Address'), hdr.index('Source')
iie, ite, ism = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = int(body[0][ia], 16)
agg = defaultdict(lambda: [0, 0, 0, 0])
tot_i = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    key = line_of.get(off, '?')
    ie, te, sm = int(r[iie] or 0), int(r[ite] or 0), int(r[ism] or 0)
    a = agg[key]
    a[0] += ie
    a[1] += te
    a[2] += sm
    a[3] += 1
    tot_i += ie
    tot_s += sm
print(f'total warp-inst {tot_i:,}   samples {tot_s:,}   static instr {len(body)}')
srcs = {}
def src_line(key):
    try:
        f, n = key.split(':')
        for d in ('/root/repo/asuna_b200/csrc/',):
            if f not in srcs:
                srcs[f] = open(d + f).read().split('\n')
            return srcs[f][int(n) - 1].strip()[:100]
    except Exception:
        return ''
for key, (ie, te, sm, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100*ie/tot_i:5.1f}% inst {100*sm/max(tot_s,1):5.1f}% smp {te/max(ie,1):5.1f} thr {n:4d} sass  {key:22s} {src_line(key)}')
