"""Per-source-line hot spots from `ncu --page source --csv --print-source cuda`.
usage: python tools/ncu_src.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Name':
        fname = r[1].split('/')[-1]
        hdr = None
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and len(hdr) > 4:
        d = dict(zip(hdr, r))
        try:
            ie = int(d.get('Instructions Executed', '0') or 0)
            te = int(d.get('Thread Instructions Executed', '0') or 0)
            smp = int(d.get('# Samples', '0') or 0)
        except ValueError:
            continue
        if ie:
            out.append((ie, te, smp, fname, d['Line No'], d['Source'].strip()[:110]))
tot = sum(o[0] for o in out)
tots = sum(o[2] for o in out)
print(f'total warp-inst {tot:,}  samples {tots:,}')
for ie, te, smp, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f'{100*ie/tot:5.1f}% inst  {100*smp/max(tots,1):5.1f}% smp  {te/ie:5.1f} thr  {f}:{ln}  {src}')
