"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time within the last
complete optimisation step (between the last two chamfer_sym launches)."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r); rows = list(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [row[ki] for row in rows]; vals = [float(row[vi].replace(",", "")) for row in rows]
sym = [i for i, n in enumerate(names) if "chamfer_sym" in n]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
a, b = (sym[-1 - k], sym[-k]) if len(sym) > k else (0, len(rows))
agg = collections.OrderedDict()
for i in range(a, b):
    n = names[i].split("(")[0][:72]
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += vals[i]
tot = sum(v[1] for v in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{t/1e3:10.1f} us {c:3d}x {100*t/tot:5.1f}%  {n}")
print(f"{tot/1e3:10.1f} us total over {b-a} launches")
