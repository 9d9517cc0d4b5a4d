"""Per-kernel totals and one PCG iteration, launch by launch, from an `ncu --metrics gpu__time_duration.sum --csv` log of
bench.py (profiles/r02_grid256_launches.csv -> profiles/r02_grid256_launches_summary.txt).  python tools/launch_summary.py <csv>"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_k, i_v, i_grid, i_blk = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")


def short(k):
    m = re.match(r"([A-Za-z_0-9]+)(<[^>]*>)?", k.replace("void ", ""))
    return m.group(1) + (m.group(2) or "")


L = [(short(r[i_k]), float(r[i_v].replace(",", "")) / 1e3, r[i_grid], r[i_blk]) for r in rows[1:]]
names = [x[0] for x in L]
a = names.index("k_assemble") if "k_assemble" in names else 0
tot, cnt = collections.OrderedDict(), collections.Counter()
for n, t, g, b in L:
    tot[n] = tot.get(n, 0.0) + t
    cnt[n] += 1
T = sum(tot.values())
print("%d consecutive launches (k_assemble, the start of a step, at launch %d)\n" % (len(L), a))
print("%-28s %8s %12s %10s %7s" % ("kernel", "launches", "total us", "avg us", "share"))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-28s %8d %12.1f %10.2f %6.1f%%" % (n, cnt[n], t, t / cnt[n], 100 * t / T))
print("%-28s %8d %12.1f" % ("all", len(L), T))
spmv = [i for i, n in enumerate(names) if n.startswith("k_pcg_spmv") and i > a + 60]
s0, s1 = spmv[2], spmv[3]
print("\none PCG iteration, launch by launch (launches %d..%d):" % (s0, s1 - 1))
print("%-28s %10s %14s %14s" % ("kernel", "us", "grid", "block"))
for k in range(s0, s1):
    print("%-28s %10.2f %14s %14s" % L[k])
seq = list(range(s0, s1))
down0 = [k for k in seq if L[k][0].startswith("k_mg_down")][0]
up0 = [k for k in seq if L[k][0].startswith("k_mg_up")][-1]
coarse = sum(L[k][1] for k in seq if down0 < k < up0)
it = sum(L[k][1] for k in seq)
print("iteration total %.1f us; level-0 down %.1f, level-0 up %.1f; coarse levels (everything between them) %.1f us = %.1f %%"
      % (it, L[down0][1], L[up0][1], coarse, 100 * coarse / it))
