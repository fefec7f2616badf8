"""Print a window of the per-instruction page of an ncu report exported by tools/ncu_box.sh.
usage: ncu_src.py <name.source.csv.gz> [first [last]] | --top N"""
import csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
h, data = rows[1], rows[2:]
ci = {c: i for i, c in enumerate(h)}
tot = sum(int(r[ci['# Samples']]) for r in data)
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
def fmt(i, r):
    s = int(r[ci['# Samples']])
    top = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stalls), reverse=True)[:2]
    tops = ' '.join(f"{n}:{v}" for v, n in top if v)
    return f"#{i:5d} ex {int(r[ci['Instructions Executed']]):>10d} smp {100*s/tot:5.2f}%  {r[ci['Source']].strip()[:86]:86s} {tops}"
if len(sys.argv) > 2 and sys.argv[2] == '--top':
    n = int(sys.argv[3])
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][ci['# Samples']]))[:n]:
        print(fmt(i, data[i]))
else:
    a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    b = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
    for i in range(a, min(b, len(data))):
        print(fmt(i, data[i]))
