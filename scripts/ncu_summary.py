"""`ncu -i X.ncu-rep --page raw --csv` -> the handful of metrics DESIGN.md / profiles/ quote, one column per launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
want = sys.argv[2:] or [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]
print("| metric | unit | " + " | ".join("launch %s" % r[0] for r in rows[2:]) + " |")
print("|---|---|" + "---|" * (len(rows) - 2))
for k in H:
    if any(w == k or (w.endswith("*") and w[:-1] in k) for w in want):
        i = H.index(k)
        print("| `%s` | %s | " % (k, U[i]) + " | ".join(r[i] for r in rows[2:]) + " |")
names = H.index("Kernel Name")
print("\nkernels:", [r[names][:60] for r in rows[2:]])
