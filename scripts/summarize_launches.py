"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel share table (markdown)."""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("vaeassoc::<unnamed>::", "").replace("vaeassoc::(anonymous namespace)::", "")
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k, n, t, 100 * t / tot))
print("\ntotal %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
