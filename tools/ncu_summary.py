"""Summarise an ncu report into the tables kept under profiles/ (run where ncu is installed,
no GPU needed):  python tools/ncu_summary.py gpurun_out/r01_full.ncu-rep gpurun_out/r01_launches.csv"""
import collections
import csv
import json
import subprocess
import sys

rep, launches = sys.argv[1], sys.argv[2]
WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("derived__lts__lts2xbar_bytes.sum.per_second", "l2_to_sm_TBps"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
kernels = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    rec = {}
    for metric, short in WANT:
        if metric in col:
            v = r[col[metric]].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            u = units[col[metric]]
            if short.endswith("_MB") and u == "Kbyte":
                v /= 1e3
            if short.endswith("_MB") and u == "byte":
                v /= 1e6
            if short == "time_us" and u == "ns":
                v /= 1e3
            if short == "time_us" and u == "ms":
                v *= 1e3
            if short == "l2_to_sm_TBps":
                v = {"Tbyte": v, "Gbyte": v / 1e3, "Mbyte": v / 1e6}.get(u.split("/")[0], v)
            rec[short] = v
    kernels.setdefault(name, rec)          # first launch of each kernel
# launch list: mean duration per kernel of the plain metric-only pass
lr = list(csv.reader(l for l in open(launches) if l.startswith('"')))
ki, vi = lr[0].index("Kernel Name"), lr[0].index("Metric Value")
dur = collections.OrderedDict()
for r in lr[1:]:
    dur.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")) / 1e3)
# bench.py --steps 3 --warmup 3 launches every kernel 6 times at full batch before its
# end-to-end (chunked, small-batch) and per-stage passes: average those 6 only
FIRST = 6
json.dump({"kernels": kernels, "launch_us": {k: sum(v[:FIRST]) / len(v[:FIRST]) for k, v in dur.items()},
           "launches_averaged": FIRST}, sys.stdout, indent=1)
