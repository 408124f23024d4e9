"""Compact per-launch summary of an ncu --set full report (read with `ncu -i ... --page raw --csv`): duration, lanes per
instruction (warp-execution efficiency), issue-active, resident warps, L1 / L2 hit rates, L1 data-pipe utilisation, DRAM bytes
and DRAM throughput, the dominant stall reasons. usage: ncu_summary.py report.ncu-rep [title] > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}


def g(r, k, default=float("nan")):
    try:
        return float(r[idx[k]].replace(",", ""))
    except Exception:
        return default


print(title)
print("source:", rep, "(ncu --set full --clock-control none; per-launch values; cold-cache, serialised replays)")
print()
print("%-3s %-34s %9s %6s %6s %6s %6s %6s %6s %9s %9s %7s %5s  %s" % ("#", "kernel", "time us", "lanes", "issue%", "warps%", "L1hit%", "L2hit%", "L1pipe%", "dram rd MB", "dram wr MB", "dram GB/s", "regs", "top stalls (warps per issue)"))
for n, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0][:34]
    t = g(r, "gpu__time_duration.sum")
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    ur, uw = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    rd *= scale.get(ur, 1.0)
    wr *= scale.get(uw, 1.0)
    st = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except Exception:
                pass
    st = [s for s in sorted(st, reverse=True) if s[1] not in ("selected",)][:3]
    print("%-3d %-34s %9.1f %6.2f %6.1f %6.1f %6.1f %6.1f %6.1f %9.1f %9.1f %7.0f %5d  %s" % (
        n, name, t, g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "l1tex__t_sector_hit_rate.pct"), g(r, "lts__t_sector_hit_rate.pct"),
        g(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), rd, wr, (rd + wr) / max(t, 1e-9) * 1e3,
        int(g(r, "launch__registers_per_thread", 0)), ", ".join("%s %.1f" % (b, a) for a, b in st)))
