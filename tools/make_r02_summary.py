"""profiles/r02_summary.md from the files tools/collect_profiles.sh produced (copied under profiles/).
usage: python tools/make_r02_summary.py"""
import json

P = "profiles/"
line = lambda f: json.load(open(P + f))
b = line("r02_bench.json")
d = line("r02_kernels.json")
k, L = d["kernels"], d["launch_us"]
tj = line("traffic.json")
tot = sum(v for kk, v in L.items() if kk.startswith("ls_") and "refresh" not in kk)
rows = []
for kk, v in sorted(L.items(), key=lambda kv: -kv[1]):
    if not kk.startswith("ls_") or "refresh" in kk:
        continue
    m = k.get(kk, {})
    l2 = ("%.2f" % m["l2_to_sm_TBps"]) if isinstance(m.get("l2_to_sm_TBps"), float) else "-"
    rows.append("| `%s` | %.1f | %.1f%% | %.1f / %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %d | %.1f / %.1f / %.1f |" % (
        kk.split("<")[0], v, 100 * v / tot, m.get("dram_read_MB", 0), m.get("dram_write_MB", 0), m.get("dram_pct", 0),
        m.get("l2_hit_pct", 0), m.get("l1_hit_pct", 0), l2, m.get("issue_active_pct", 0), m.get("warps_active_pct", 0),
        int(m.get("regs", 0)), m.get("stall_long_sb", 0), m.get("stall_barrier", 0), m.get("stall_short_sb", 0)))
table = "\n".join(rows)
n2, n4, n8 = [line("r02_bench_%s.json" % x) for x in ("n2", "n4", "n8")]
t8, t1, tr = line("r02_bench_train8.json"), line("r02_bench_train.json"), line("r02_bench_train_reference.json")
ag, agr = line("r02_bench_agent.json")["latency"], line("r02_bench_agent_reference.json")["latency"]
ref = line("r02_bench_reference.json")
dom = b["roofline"]["kernel"]
traffic = b["roofline"]["traffic"] or tj["cfg2/fp32/channels_last"][dom]
bwd_traffic = tj["cfg2/fp32/channels_last"]["splat_bwd(transpose+gather)"]
md = f'''# Round 02 profile summary (B200, sm_100a)

Command profiled: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-graph`
(kernel-by-kernel launches of the same step the bench replays as a CUDA graph; BASELINE.json configs[1]: fwd+bwd, B=16,
4 cams, D=48, C=64, 200x200, fp32; default layouts: channels-last BEV + gradient, NCHW features).
Everything here was produced by `tools/collect_profiles.sh` on one fresh B200 box (multi-GPU lines: separate `gpurun --gpus N`
calls of the same commit series), tables by `tools/make_r02_summary.py`.

* `r02_launches.csv` - every launch with `gpu__time_duration.sum` (`ncu --metrics gpu__time_duration.sum --clock-control none`);
  cold-cache, serialised: compare SHARES, not absolutes.  The first 6 launches of each kernel are the full-batch steps.
* `r02_kernels.json` - per-kernel metrics from `ncu --set full --clock-control none --import-source on` (first launch of each
  kernel) + mean duration of the 6 full-batch launches (`tools/ncu_summary.py`).
* `r02_timeline.txt` - CUPTI timeline of the graph replay (`tools/timeline.py graph`): in-situ durations, idle time.
* `traffic.json` - DRAM read+write bytes per launch of the dominant stages (read by `bench.py` for `roofline.traffic`).
* `r02_bench*.json` - bench lines of the same build, all taken WITHOUT a profiler (see the table below).

Bench (no profiler): **{b["ms_per_step"]:.4f} ms/step, {b["value"]:.0f} samples/s**, step roofline {b["roofline_step"]["frac"]:.3f} of the measured
HBM peak ({b["roofline"]["peak"]:.1f} GB/s); dominant stage `{dom}` ({b["roofline"]["kernels"]}) at
{b["roofline"]["frac"]:.3f}, its DRAM traffic {traffic / 1e6:.0f} MB for {b["roofline"]["algorithmic_bytes_per_launch"] / 1e6:.0f} MB algorithmic;
e2e (host buffers) {b["e2e"]["value"]:.0f} samples/s = {b["e2e"]["ms_per_step"]:.2f} ms against a measured link floor of {b["e2e"]["link_floor_ms"]:.2f} ms for the same bytes.
Round 1: 0.3144 ms/step, step roofline 0.205, dominant stage 0.28 with 450 MB of traffic (2.02x algorithmic).

| kernel | launch (us) | share | DRAM read / write (MB) | DRAM % | L2 hit % | L1 hit % | L2->SM (TB/s) | issue active % | warps active % | regs | stalls long_sb / barrier / short_sb |
|---|---|---|---|---|---|---|---|---|---|---|---|
{table}

Sum of `ls_*` kernels per step (ncu, serialised, cold caches): {tot:.1f} us.  In situ (CUPTI, graph replay, `r02_timeline.txt`): span
244.9 us per step, idle 2.3 us; index 26.7, place 26.9, canon 20.9, splat 58.8, gather 69.8, softmax backward 17.0, camera 3.9,
zero 3.8, scan 6.2; softmax 8.1 / NHWC staging 11.7 / 8.4 run on side streams.  Backward = gather + softmax backward = 86.8 us
(round 1: 137 us with the transposer).

## All bench lines of this build

| file | configuration | ms/step | samples/s | roofline (dominant / step) |
|---|---|---|---|---|
| `r02_bench.json` | default (channels-last BEV, NCHW features, fp32) | {b["ms_per_step"]:.4f} | {b["value"]:.0f} | {b["roofline"]["frac"]:.3f} / {b["roofline_step"]["frac"]:.3f} |
'''
for f, desc in (("r02_bench_featcl.json", "+ channels-last features (LS_FEAT_NHWC: no staging copies)"),
                ("r02_bench_bf16.json", "bf16 features / logits"),
                ("r02_bench_bf16_bev.json", "bf16 features / logits + opt-in bf16 BEV and gradient (`--bev-dtype bf16`)"),
                ("r02_bench_nchw.json", "NCHW BEV + gradient (the reference's strides; staged backward)"),
                ("r02_bench_bulk_tma.json", "`LS_SPLAT_OUT=bulk`: one bulk (TMA) store per tile instead of direct rows"),
                ("r02_bench_stress.json", "stress: B=32, 6 cams, D=96, 400x400 (configs[3]); round 1: 2.01 ms"),
                ("r02_bench_stress_bf16.json", "stress, bf16")):
    x = line(f)
    md += f'| `{f}` | {desc} | {x["ms_per_step"]:.4f} | {x["value"]:.0f} | {x["roofline"]["frac"]:.3f} / {x["roofline_step"]["frac"]:.3f} |\n'
sc = b["static_rig_cache"]
md += f'''| `r02_bench.json` key `static_rig_cache` | opt-in static-rig cache (index structures reused, weights refreshed; bit-identical) | {sc["ms_per_step"]:.4f} | {sc["value"]:.0f} | - |
| `r02_bench.json` key `gpu_reference` | the reference's torch op chain on the same GPU, full batch | {b["gpu_reference"]["ms_per_step"]:.0f} | {b["gpu_reference"]["value"]:.1f} | - |
| `r02_bench_reference.json` | `--impl reference`: the same on {ref["cpu_baseline"]["cores"]} host threads, FULL batch of 16 | {ref["ms_per_step"]:.0f} | {ref["value"]:.2f} | - |

Multi-GPU (torchrun, one rank per GPU, per-GPU batch 16 for the lift-splat and 12 for training; `r02_bench_n2/n4/n8.json`,
`r02_bench_train.json`, `r02_bench_train8.json`, `r02_nccl_n8.txt`):

| GPUs | lift-splat samples/s (device) | ms/step | efficiency | e2e samples/s (ms) | train samples/s (ms/step) | train efficiency |
|---|---|---|---|---|---|---|
| 1 | {b["value"]:.0f} | {b["ms_per_step"]:.4f} | 1.00 | {b["e2e"]["value"]:.0f} ({b["e2e"]["ms_per_step"]:.2f}) | {b["train"]["samples_per_s"]:.1f} ({b["train"]["ms_per_step"]:.1f}) | 1.00 |
'''
for n, x in ((2, n2), (4, n4), (8, n8)):
    md += (f'| {n} | {x["value"]:.0f} | {x["ms_per_step"]:.4f} | {x["value"] / n / b["value"]:.3f} | {x["e2e"]["value"]:.0f} ({x["e2e"]["ms_per_step"]:.2f}) | '
           f'{x["train"]["samples_per_s"]:.1f} ({x["train"]["ms_per_step"]:.1f}) | {x["train"]["samples_per_s"] / n / b["train"]["samples_per_s"]:.3f} |\n')
md += f'''| 8 (`--workload train --steps 20`) | - | - | - | - | {t8["value"]:.1f} ({t8["ms_per_step"]:.1f}) | {t8["value"] / 8 / t1["value"]:.3f} |

* The lift-splat has no data-path collective: device-timed efficiency is 1.00 by construction.
* Training: DDP all-reduces 78.4 MB of gradients per step (19.6 M parameters) inside the timed region; the step grows from
  52.7 ms (1 GPU) to 57-60 ms (8 GPUs): efficiency 0.88 (8 timed steps inside the default line) to 0.92 (20-step run),
  above the 0.85 target.  NCCL reports no NVLS on these VMs (`r02_nccl_n8.txt`); what is lost is the exposed tail of the
  all-reduce after the last bucket (the camera-encoder trunk's gradients are produced last) plus per-rank jitter of the
  3 000+ small kernels of the stock PyTorch stack; the lift-splat library is 0.5 % of the step.
* With the reference's own lift-splat ops in the same stack the 1-GPU step is {tr["ms_per_step"]:.0f} ms ({tr["value"]:.1f} samples/s):
  the library makes the full training step {tr["ms_per_step"] / t1["ms_per_step"]:.1f}x faster.
* e2e does not scale past 2 GPUs on these hosts: 4.7 / 5.3 / 16.0 / 25.1 ms per step at 1 / 2 / 4 / 8 ranks for 2 x 206 MB per
  rank.  The hosts are single-NUMA VMs (`numa_node` = -1 for every GPU, so the NUMA binding in `bench.py` is a no-op);
  `e2e.link_floor_ms` (the same bytes copied both ways by all ranks at once, no kernels) is what PCIe plus the host memory
  system allow - at 1 GPU the step is within 5 % of it.

Agent tick (B=1, `r02_bench_agent.json`, 1000 iterations): as one CUDA graph p50 {ag["graph"]["wall_ms"]["p50"]:.2f} ms / p99 {ag["graph"]["wall_ms"]["p99"]:.2f} ms
wall ({ag["graph"]["device_ms"]["p50"]:.2f} / {ag["graph"]["device_ms"]["p99"]:.2f} device); stream launches {ag["stream"]["wall_ms"]["p50"]:.1f} / {ag["stream"]["wall_ms"]["p99"]:.1f} ms; with the reference's torch
lift-splat ops in the same stack {agr["stream"]["wall_ms"]["p50"]:.1f} / {agr["stream"]["wall_ms"]["p99"]:.1f} ms (its host syncs prevent graph capture).  Paper: 74.92 ms on a Quadro RTX 5000.

## Reading

* **The gradient round trip is gone.**  Round 1 staged the NCHW gradient as cell rows (`ls_bwd_transpose_kernel`, 54 us,
  170 MB read + 87 MB written) and gathered from the copy (66 us, another 170 MB).  With a channels-last gradient the gather
  reads the rows in place: {bwd_traffic / 1e6:.1f} MB of DRAM traffic for the whole backward stage against 222.6 MB algorithmic ({bwd_traffic / 222.56e6:.2f}x; round 1:
  450 MB, 2.02x).  `r02_nchw.ncu-rep` has the compat path of the same build: transposer 45.4 us (169.8 + 82.5 MB), gather 64.4 us
  (170.2 + 22.1 MB), NCHW splat 86.7 us.
* **The in-place gather is cold-DRAM + issue bound.**  170 MB read = 118 MB of distinct gradient rows + feature rows + the
  pixel-major index: essentially every byte once; L1 hit 28 % (a CTA is one feature-map column, its rays share cells), L2 hit
  44 %, 56 % issue-active, 24 warps/SM at 80 registers.  A predicated row load for dropped points halved its speed (ptxas
  serialised the window); selecting the address of a zero row keeps all eight loads in flight.
* **Forward splat: three write-outs, same records.**  NCHW tile 86.7 us, channels-last bulk (TMA) store 83.1 us
  (`r02_bulk_tma.ncu-rep`: 43.4 + 107.8 MB, L1 hit 10 %), direct row stores 63.6 us (41.4 + 106.2 MB, L1 hit 16.5 %, 7 CTAs/SM at 72
  registers, 640 B of shared memory).  The kernel sits at 63 % of the L1 data-pipe wavefront peak and 53 % issue-active; L2->SM
  9.4 TB/s.  It is NOT L2-read bound: square tiles that cut L2 reads in principle (3.6 records per pixel and tile instead of
  1.15) were 8-11 us slower, fewer or more rows in flight per quarter-warp did not help, packed FFMA2 was slower.
* **The integer pipeline is now a third of the step** (zero 3.8 + index 26.7 + scan 6.2 + place 26.9 + canon 20.9 = 84.5 us):
  the index kernel is ~150 exact float32 instructions per point with two IEEE divisions (the third, for z, is proven away),
  placement is bound by its scattered 8-byte stores, canon by k^2 key compares.  The opt-in static-rig cache replaces all five by
  two streaming refresh kernels (17.5 + 10.4 us under ncu): 0.195 ms per step.
* SASS: `UBLKCP.G.S` (bulk async copy shared -> global, the TMA path) is in `ls_splat_fwd_kernel<.., LS_OUT_NHWC_BULK, 64>`
  (`cuobjdump -sass libls_b200.so | grep UBLKCP`: fp32 and bf16 variants).
'''
open(P + "r02_summary.md", "w").write(md)
print("written", len(md))
