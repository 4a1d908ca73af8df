"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics that matter here."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
cols = [("dur_us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("inst_M", "smsp__inst_executed.sum"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smem%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed")]
print("%-28s %-14s " % ("kernel", "grid") + " ".join("%8s" % c[0] for c in cols))
for d in data:
    name = d[idx["Kernel Name"]].replace("void ", "").split("(")[0][:28]
    vals = []
    for c, m in cols:
        if m not in idx:
            vals.append("     n/a"); continue
        try:
            v = float(d[idx[m]].replace(",", ""))
        except ValueError:
            vals.append("     n/a"); continue
        if c == "inst_M":
            v /= 1e6
        elif c.endswith("_MB"):
            v *= TO_MB.get(units[idx[m]], 1.0)
        elif c == "dur_us":
            v *= TO_US.get(units[idx[m]], 1.0)
        vals.append("%8.2f" % v)
    print("%-28s %-14s " % (name, d[idx["Grid Size"]].replace(" ", "")) + " ".join(vals))
