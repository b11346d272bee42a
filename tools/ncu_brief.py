"""Print the handful of ncu metrics used while tuning (run where ncu is installed, no GPU):
python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "derived__lts__lts2xbar_bytes.sum.per_second", "lts__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("=====", r[hdr.index("Kernel Name")][:70])
    for w in want:
        if w in hdr:
            print("  %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = sorted(((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
        "_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:6]
    print("  stalls:", ", ".join("%s %.2f" % (n, v) for v, n in st))
