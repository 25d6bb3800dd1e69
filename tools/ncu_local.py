"""Local-memory traffic of one kernel by source line: joins an `ncu --page source --csv` SASS dump with `nvdisasm -g -c`
line info and sums, per line, the executed STL / LDL warp instructions, their active threads and the L2 sectors they
imply.  usage: python tools/ncu_local.py <ncu_source.csv> <nvdisasm.txt> <kernel substring> [top_n] [kernel index]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
line_of, cur, active = {}, None, False
for ln in open(dis):
    if (ln.startswith('.text.') and ln.rstrip().endswith(':')) or (ln.startswith('//-----') and '.text.' in ln):
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
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[starts[which]]
body = []
for r in rows[starts[which] + 1:]:
    if not r or r[0] in ('Kernel Name', 'Address'):
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
base = int(body[0][col['Address']], 16)
agg = defaultdict(lambda: [0, 0, 0, 0])  # st inst, st threads, ld inst, sectors
tot = [0, 0, 0, 0, 0]
for r in body:
    ie, te = int(r[col['Instructions Executed']] or 0), int(r[col['Thread Instructions Executed']] or 0)
    tot[4] += ie
    s = r[col['Source']]
    if 'STL' not in s and 'LDL' not in s:
        continue
    a = agg[line_of.get(int(r[col['Address']], 16) - base, '?')]
    sec = int(r[col['L2 Theoretical Sectors Local']] or 0)
    if 'STL' in s:
        a[0] += ie; a[1] += te; tot[0] += ie; tot[1] += te
    else:
        a[2] += ie; tot[2] += ie
    a[3] += sec; tot[3] += sec
print(f'warp instructions {tot[4]:,}; STL {tot[0]:,} ({tot[1] / max(tot[0], 1):.1f} threads each), LDL {tot[2]:,}, theoretical L2 sectors local {tot[3]:,} ({tot[3] * 32 / 1e9:.2f} GB)')
for k, a in sorted(agg.items(), key=lambda x: -x[1][3])[:top]:
    print(f'{a[3] * 32 / 1e6:9.1f} MB  STL {a[0]:>10,} x{a[1] / max(a[0], 1):5.1f}thr  LDL {a[2]:>10,}  {k}')
