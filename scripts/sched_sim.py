"""Discrete-event model of one launch of the persistent tile kernel (csrc/gemm_group.cu), calibrated with the measured
per-task phases (profiles/r1_group_final_ncu.md), to compare TASK ORDERS of a fused segment offline.

Model per CTA pair ("cluster"): tasks are popped from one global queue in list order; the TMA producer may run ahead of the
MMA by the stage ring, so a task's main loop starts at max(previous main loop end, dependencies visible + first-operand
latency, accumulator of two tasks ago drained); its epilogue starts at max(main loop end, previous epilogue end) and, if
the tile has dependents, ends with the release (+signal).  Durations: k-block 0.36 us (256-wide tile; 0.2 us for tiles
<= 64 wide), epilogue 1.38 us per 32-column chunk per warp (4 chunks for a 256-wide tile), signal 0.9 us.
"""
import heapq, sys

T_KB_WIDE, T_KB_NARROW = 0.36, 0.20
T_CHUNK, T_SIGNAL, T_FIRST = 1.38, 0.9, 1.5
T_REDUCE_CHUNK = 0.65
N_CLUSTERS = 74


class Task(object):
    __slots__ = ("name", "nkb", "bn", "deps", "reduce", "signals", "done", "mod", "layer", "rb")

    def __init__(self, name, nkb, bn, deps, reduce=False, signals=True, mod=0, layer=0, rb=0):
        self.name, self.nkb, self.bn, self.deps, self.reduce, self.signals = name, nkb, bn, deps, reduce, signals
        self.done = None; self.mod, self.layer, self.rb = mod, layer, rb


def simulate(tasks, n_clusters=N_CLUSTERS, start=1.2):
    """tasks: list in queue order; deps = list of Task objects.  Returns (makespan, per-task times)."""
    # clusters pop in order; event-driven: each cluster has state (mma_free, epi_free, acc_free[2], count)
    cl = [dict(mma=start, epi=start, acc=[start, start], n=0, pop=start) for _ in range(n_clusters)]
    heap = [(start, i) for i in range(n_clusters)]     # time at which cluster can pop its next task
    heapq.heapify(heap)
    end = 0.0
    for t in tasks:
        pop_time, ci = heapq.heappop(heap)
        c = cl[ci]
        # dependencies: their `done` must be known -- list order guarantees they were scheduled earlier
        dep_ready = max([d.done for d in t.deps], default=0.0)
        ops_ready = max(pop_time, dep_ready) + (T_FIRST if dep_ready > c["mma"] - T_FIRST else 0.0)
        acc = c["n"] & 1
        t_kb = T_KB_WIDE if t.bn > 64 else T_KB_NARROW
        mma_start = max(c["mma"], ops_ready if t.deps else max(pop_time + (T_FIRST if c["n"] == 0 else 0.0), c["mma"]), c["acc"][acc])
        mma_end = mma_start + t.nkb * t_kb
        chunks = max(1, min(t.bn // 32, 8)) / 2.0
        epi_start = max(mma_end, c["epi"])
        epi_end = epi_start + chunks * (T_REDUCE_CHUNK if t.reduce else T_CHUNK) + (T_SIGNAL if t.signals else 0.0)
        t.done = epi_end
        c["mma"], c["epi"] = mma_end, epi_end
        c["acc"][acc] = epi_end - (T_SIGNAL if t.signals else 0.0)
        c["n"] += 1
        # the scheduler pops the next task when this one's first stages are in flight
        heapq.heappush(heap, (mma_start, ci))
        end = max(end, epi_end)
    return end


def seg_fwd(B=8192, order="layer", dims=((784, 500, 500, 8), (147, 200, 200, 8))):
    """encoder forward: enc1 -> enc2 -> heads, both modalities"""
    RB = (B + 255) // 256
    def tn(n): return (n + 255) // 256
    def bn(n):
        t = tn(n); return min(256, ((n + t - 1) // t + 63) // 64 * 64)
    tasks = {}
    for m, (k0, n1, n2, n3) in enumerate(dims):
        ks = [k0, n1, n2]; ns = [n1, n2, n3]
        for rb in range(RB):
            prev = []
            for L in range(3):
                cur = [Task("m%d.L%d.rb%d.n%d" % (m, L, rb, j), (ks[L] + 31) // 32, bn(ns[L]), list(prev), signals=(L < 2), mod=m, layer=L, rb=rb)
                       for j in range(tn(ns[L]))]
                tasks[(m, L, rb)] = cur
                prev = cur
    M = len(dims)
    out = []
    if order == "layer":                      # current: layer-major, modality, row block
        for L in range(3):
            for m in range(M):
                for rb in range(RB):
                    out += tasks[(m, L, rb)]
    elif order == "layer_rb":                 # layer-major, row block, modality (modalities interleaved)
        for L in range(3):
            for rb in range(RB):
                for m in range(M):
                    out += tasks[(m, L, rb)]
    elif order.startswith("wave"):            # row blocks in G groups; group g runs one layer behind group g-1
        G = int(order[4:])
        per = (RB + G - 1) // G
        groups = [list(range(g * per, min(RB, (g + 1) * per))) for g in range(G)]
        for step in range(3 + G - 1):
            for g in range(G):
                L = step - g
                if 0 <= L < 3:
                    for rb in groups[g]:
                        for m in range(M):
                            out += tasks[(m, L, rb)]
    elif order == "small_first":              # the short joint-modality chain first, then the image layers
        for L in range(3):
            for rb in range(RB):
                out += tasks[(1, L, rb)]
        for L in range(3):
            for rb in range(RB):
                out += tasks[(0, L, rb)]
    elif order == "img_then_fill":            # image layer L, then joint layer L (fills the bubble of the image boundary)
        for L in range(3):
            for rb in range(RB):
                out += tasks[(0, L, rb)]
            for rb in range(RB):
                out += tasks[(1, L, rb)]
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    for order in ("layer", "layer_rb", "img_then_fill", "small_first", "wave2", "wave4", "wave8"):
        ts = seg_fwd(B, order)
        print("%-14s %4d tasks  makespan %.1f us" % (order, len(ts), simulate(ts)))
