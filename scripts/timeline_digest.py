"""Digest of a VAEASSOC_TC_TIMELINE_ALL dump of ONE launch: per problem (layer) the task count, start / end of its main
loops and epilogues, mean durations; per cluster the share of time its MMA issuer was busy; the launch's critical tail."""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
pat = re.compile(r"task\s+(\d+) prob\s+(\d+) \(\s*(\d+),\s*(\d+)\) nkb\s+(\d+) cl\s+(\d+) entry\s+([\d.-]+) \| prod\s+([\d.-]+) mma\s+([\d.-]+)\.\.\s*([\d.-]+) \((\d+) cyc\) epi\s+([\d.-]+)\.\.\s*([\d.-]+) us")
heads = [i for i, l in enumerate(lines) if l.startswith("[group timeline]")]
lines = lines[heads[-1]:]          # the last dump of the file (bench.py profiles the step several times)
T = []
for l in lines:
    m = pat.search(l)
    if m: T.append([float(x) for x in m.groups()])
print(next(l for l in lines if l.startswith("[group timeline]")))
end = max(t[12] for t in T)
byp = collections.defaultdict(list)
for t in T: byp[int(t[1])].append(t)
print("prob  n  nkb | prod first..last | mma: first start, last end, mean dur, cyc/kb | epi: mean dur, last end | mean wait prod->mma-start")
for p, ts in sorted(byp.items()):
    print("%4d %4d %4d | %6.1f..%6.1f | %6.1f %6.1f %5.2f %5.0f | %5.2f %6.1f | %5.2f" % (
        p, len(ts), ts[0][4], min(t[7] for t in ts), max(t[7] for t in ts), min(t[8] for t in ts), max(t[9] for t in ts),
        sum(t[9] - t[8] for t in ts) / len(ts), sum(t[10] / max(t[4], 1) for t in ts) / len(ts),
        sum(t[12] - t[11] for t in ts) / len(ts), max(t[12] for t in ts), sum(max(0, t[8] - t[7]) for t in ts) / len(ts)))
bycl = collections.defaultdict(list)
for t in T: bycl[int(t[5])].append(t)
busy = [sum(t[9] - t[8] for t in ts) / end for ts in bycl.values()]
epi = [sum(t[12] - t[11] for t in ts) / end for ts in bycl.values()]
print("clusters %d: MMA-issuer busy share mean %.3f min %.3f max %.3f ; epilogue busy share mean %.3f" % (len(bycl), sum(busy) / len(busy), min(busy), max(busy), sum(epi) / len(epi)))
# time buckets: how many clusters have an MMA main loop in flight
nb = 30
for b in range(nb):
    t0, t1 = end * b / nb, end * (b + 1) / nb
    act = sum(max(0.0, min(t[9], t1) - max(t[8], t0)) for t in T) / (t1 - t0)
    ep = sum(max(0.0, min(t[12], t1) - max(t[11], t0)) for t in T) / (t1 - t0)
    probs = collections.Counter(int(t[1]) for t in T if t[8] < t1 and t[9] > t0)
    print("%6.1f-%6.1f us: main loops in flight %5.1f  epilogues %5.1f  probs %s" % (t0, t1, act, ep, dict(probs.most_common(5))))
