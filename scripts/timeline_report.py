"""Per-problem / per-cluster digest of the VAEASSOC_TC_TIMELINE_ALL dump (stderr of bench.py / debug_gemm)."""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith("[group timeline]")]
which = [int(a) for a in sys.argv[2:]] or list(range(4))
pat = re.compile(r"task\s+(\d+) prob\s+(\d+) \(\s*(\d+),\s*(\d+)\) nkb\s+(\d+) cl\s+(\d+) entry\s+([\d.-]+) \| prod\s+([\d.-]+) mma\s+([\d.-]+)\.\.\s*([\d.-]+) \((\d+) cyc\) epi\s+([\d.-]+)\.\.\s*([\d.-]+) us(?: \| ld0\s+([\d.-]+) chunk0\s+([\d.-]+) loop\s+([\d.-]+))?(?: \| math\s+([\d.-]+) mask\s+([\d.-]+) wait\s+([\d.-]+) sts\s+([\d.-]+))?")
for w in which:
    a = heads[w]; b = heads[w + 1] if w + 1 < len(heads) else len(lines)
    print(lines[a])
    T = [[float(x) if x is not None else 0.0 for x in m.groups()] for m in map(pat.search, lines[a:b]) if m]
    byprob = collections.defaultdict(list)
    for t in T: byprob[int(t[1])].append(t)
    for p, ts in sorted(byprob.items()):
        print("  prob %2d n %3d nkb %3d | prod start %.1f..%.1f | mma end max %.1f | epi dur avg %.2f | epi end max %.1f | run cyc/kb %.0f | ld0 +%.0f cyc chunk0 +%.0f cyc loop +%.2f us | math +%.0f mask +%.0f wait +%.0f sts +%.0f cyc" % (
            p, len(ts), ts[0][4], min(t[7] for t in ts), max(t[7] for t in ts), max(t[9] for t in ts),
            sum(t[12] - t[11] for t in ts) / len(ts), max(t[12] for t in ts),
            sum((t[9] - max(t[7], t[8])) * 1965 / max(t[4], 1) for t in ts) / len(ts),
            sum(t[13] for t in ts) / len(ts), sum(t[14] for t in ts) / len(ts), sum(t[15] - t[11] for t in ts) / len(ts),
            *[sum(t[k] for t in ts) / len(ts) for k in (16, 17, 18, 19)]))
