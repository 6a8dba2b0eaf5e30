"""ncu -i X.ncu-rep --page raw --csv  ->  compact per-kernel markdown table (for profiles/)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"), ("launch__registers_per_thread", "regs")]
agg = collections.OrderedDict()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    key = (r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), r[idx["Grid Size"]], r[idx["Block Size"]],
           round(float(r[idx["dram__bytes_read.sum"]].replace(",", "")) / 50) )
    agg.setdefault(key, []).append(r)
print("| kernel | grid | block | n | " + " | ".join(f"{n} ({units[idx[m]]})" for m, n in M) + " |")
print("|---|---|---|---|" + "---|" * len(M))
for (k, g, b, _), rs in agg.items():
    vals = []
    for m, n in M:
        v = [float(r[idx[m]].replace(",", "")) for r in rs]
        vals.append("%.3g" % (sum(v) / len(v)))
    print(f"| {k} | {g} | {b} | {len(rs)} | " + " | ".join(vals) + " |")
