"""Aggregate an `ncu --page source --csv --print-source sass` dump: per-opcode and per-region counts."""
import csv, collections, sys
path = sys.argv[1]; years = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('Instructions Executed')].isdigit()]
first_kernel_end = next((k for k in range(1, len(data)) if data[k][0] < data[k-1][0]), len(data)); data = data[:first_kernel_end]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); it = hdr.index('Thread Instructions Executed'); ismp = hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data); tott = sum(int(r[it]) for r in data)
print("total warp inst", tot, "thread inst", tott, "per year", tot / years, tott / years, "n sass", len(data))
byop = collections.Counter(); bythr = collections.Counter()
for r in data:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    byop[op] += int(r[ia]); bythr[op] += int(r[it])
for op, c in byop.most_common(28):
    print(f"{op:12s} {c/years:9.1f} warp-inst/yr {c/tot*100:5.1f}%  avgthr {bythr[op]/max(c,1):.1f}")
cnts = [int(r[ia]) for r in data]
segs = []; i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(cnts[j + 1] - cnts[i]) <= 0.03 * max(cnts[i], 1): j += 1
    segs.append((i, j, cnts[i])); i = j + 1
print()
for (i, j, c) in segs:
    s = sum(cnts[k] for k in range(i, j + 1))
    if s > 0.005 * tot:
        thr = sum(int(data[k][it]) for k in range(i, j + 1)) / max(1, s)
        smp = sum(int(data[k][ismp]) for k in range(i, j + 1))
        print(f"sass {i:4d}-{j:4d} n={j-i+1:4d} exec/yr={c/years:8.2f} share={s/tot*100:5.1f}% avgthr={thr:4.1f} samples={smp:6d}  {data[i][isrc][:50]}")
