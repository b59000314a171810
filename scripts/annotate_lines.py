"""Annotate the output of sass_by_line.py with the source text.  usage: annotate_lines.py <lines.txt> <main source> [min count]"""
import sys
rows = [l.split() for l in open(sys.argv[1]).read().split('\n')[1:] if l.strip()]
path = sys.argv[2]; name = path.split('/')[-1]
src = open(path).read().split('\n')
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 8
oth = {}
for r in rows:
    f, l = r[0].split(':'); l = int(l); c = float(r[1])
    if f == name:
        if c >= thr: print(f"{l:4d} {c:9.1f}  {src[l-1].strip()[:120]}")
    else:
        oth[f] = oth.get(f, 0) + c
for f, c in oth.items(): print(f"{f}: {c:.1f}")
