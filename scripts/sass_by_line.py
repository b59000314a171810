"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm --print-line-info` of the same
cubin: executed warp-instructions per CUDA source line (inlined callees attributed to their own file:line).
usage: sass_by_line.py <ncu_sass.csv> <nvdisasm.txt> <mangled kernel name> [units per launch]"""
import csv, collections, re, sys

csv_path, dis_path, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1e6
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ia].isdigit()]
end = next((k for k in range(1, len(data)) if data[k][0] < data[k - 1][0]), len(data))
data = data[:end]

lines = open(dis_path).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':'))
cur = ('?', 0); inst = []
for l in lines[start + 1:]:
    if l.startswith('//---') or l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        inst.append((cur, m.group(2)))
assert len(inst) == len(data), (len(inst), len(data))
by = collections.Counter()
for (loc, _), r in zip(inst, data):
    by[loc] += int(r[ia])
tot = sum(by.values())
print(f"total {tot / units:.1f} warp-inst per unit")
for loc, c in sorted(by.items()):
    if c / units >= 1.0:
        print(f"{loc[0]}:{loc[1]:<5d} {c / units:8.1f}  {c / tot * 100:5.1f}%")
