"""profiles/r02_summary.md + profiles/traffic.json from the files tools/collect_profiles.sh produced (copied under
profiles/).  usage: python tools/make_r02_summary.py"""
import json
import os
import re

P = "profiles/"
line = lambda f: json.load(open(P + f))
have = lambda f: os.path.exists(P + f)
b = line("r02_bench.json")
d = line("r02_kernels.json")
k, L = d["kernels"], d["launch_us"]
find = lambda kk, p: next(v for n, v in kk.items() if p in n)
mb = lambda m: (m["dram_read_MB"] + m["dram_write_MB"]) * 1e6

# ---- traffic.json: DRAM read+write bytes per launch of the two dominant stages --------------------------------------
traffic = {"source": "profiles/r02_kernels.json / r02_kernels_nchw.json (ncu --set full --clock-control none over `python "
                     "bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-graph [--bev-format "
                     "nchw]`, first launch of each kernel; DRAM read+write bytes per launch, summed over the kernels of the stage)",
           "cfg2/fp32/channels_last": {
               "splat_fwd": int(mb(find(k, "ls_canon")) + mb(find(k, "ls_splat_fwd_direct"))),
               "backward(gather+epilogue)": int(mb(find(k, "ls_bwd_gather_occ")) + mb(find(k, "ls_bwd_epilogue")))}}
if have("r02_kernels_nchw.json"):
    kn = line("r02_kernels_nchw.json")["kernels"]
    traffic["cfg2/fp32/nchw"] = {
        "splat_fwd": int(mb(find(kn, "ls_canon")) + mb(find(kn, "ls_splat_fwd_kernel"))),
        "backward(gather+epilogue)": int(mb(find(kn, "ls_bwd_transpose")) + mb(find(kn, "ls_bwd_gather_occ")) +
                                         mb(find(kn, "ls_bwd_epilogue")))}
json.dump(traffic, open(P + "traffic.json", "w"), indent=1)

# ---- per-kernel table ---------------------------------------------------------------------------------------------
tot = sum(v for kk, v in L.items() if kk.startswith("ls_") and "refresh" not in kk and kk in k)
rows = []
for kk, v in sorted(L.items(), key=lambda kv: -kv[1]):
    if not kk.startswith("ls_") or "refresh" in kk or kk not in k:      # (kernels of bench.py's stage-by-stage pass only)
        continue
    m = k.get(kk, {})
    l2 = ("%.2f" % m["l2_to_sm_TBps"]) if isinstance(m.get("l2_to_sm_TBps"), float) else "-"
    rows.append("| `%s` | %.1f | %.1f%% | %.1f / %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %d | %.1f / %.1f / %.1f |" % (
        kk.split("<")[0], v, 100 * v / tot, m.get("dram_read_MB", 0), m.get("dram_write_MB", 0), m.get("dram_pct", 0),
        m.get("l2_hit_pct", 0), m.get("l1_hit_pct", 0), l2, m.get("issue_active_pct", 0), m.get("warps_active_pct", 0),
        int(m.get("regs", 0)), m.get("stall_long_sb", 0), m.get("stall_barrier", 0), m.get("stall_short_sb", 0)))
table = "\n".join(rows)

# ---- in-situ timeline ---------------------------------------------------------------------------------------------
tl = open(P + "r02_timeline.txt").read()
span = float(re.search(r"span per step ([\d.]+)", tl).group(1))
idle = float(re.search(r"idle per step ([\d.]+)", tl).group(1))
insitu = {m.group(1).split("<")[0].strip(): float(m.group(2)) for m in re.finditer(r"^(ls_\S+).*?mean ([\d.]+) us", tl, re.M)}
t = lambda p: next(v for n, v in insitu.items() if p in n)
chain = ["camera_transform", "index", "tile_totals", "tile_scan", "place", "canon", "splat_fwd", "bwd_gather", "bwd_epilogue"]
chain_us = sum(t(p) for p in chain)

dom = b["roofline"]["kernel"]
dom_traffic = b["roofline"]["traffic"] or traffic["cfg2/fp32/channels_last"][dom]
bwd_traffic = traffic["cfg2/fp32/channels_last"]["backward(gather+epilogue)"]
fwd_traffic = traffic["cfg2/fp32/channels_last"]["splat_fwd"]
ref = line("r02_bench_reference.json")
md = f'''# Round 02 profile summary (B200, sm_100a) - final build of the round

Command profiled: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-graph`
(kernel-by-kernel launches of the same step the bench replays as a CUDA graph; BASELINE.json configs[1]: fwd+bwd, B=16,
4 cams, D=48, C=64, 200x200, fp32; default layouts: channels-last BEV + gradient, NCHW features).
Single-GPU files were produced by `tools/collect_profiles.sh` on one fresh B200 box, tables by `tools/make_r02_summary.py`.

* `r02_launches.csv` - every launch with `gpu__time_duration.sum` (`ncu --metrics gpu__time_duration.sum --clock-control none`);
  cold-cache, serialised: compare SHARES, not absolutes.  The first 6 launches of each kernel are the full-batch steps.
* `r02_kernels.json` (`r02_kernels_nchw.json`: the NCHW-compat path) - per-kernel metrics from `ncu --set full --clock-control
  none --import-source on` (first launch of each kernel) + mean duration of the 6 full-batch launches (`tools/ncu_summary.py`).
* `r02_timeline.txt` - CUPTI timeline of the graph replay (`tools/timeline.py graph`): in-situ durations, idle time.
* `traffic.json` - DRAM read+write bytes per launch of the dominant stages (read by `bench.py` for `roofline.traffic`).
* `r02_bench*.json` - bench lines of the same build, all taken WITHOUT a profiler (see the table below).

Bench (no profiler): **{b["ms_per_step"]:.4f} ms/step, {b["value"]:.0f} samples/s**, step roofline {b["roofline_step"]["frac"]:.3f} of the measured
HBM peak ({b["roofline"]["peak"]:.1f} GB/s); dominant stage `{dom}` ({b["roofline"]["kernels"]}) at
{b["roofline"]["frac"]:.3f}, its DRAM traffic {dom_traffic / 1e6:.0f} MB for {b["roofline"]["algorithmic_bytes_per_launch"] / 1e6:.0f} MB algorithmic;
e2e (host buffers) {b["e2e"]["value"]:.0f} samples/s = {b["e2e"]["ms_per_step"]:.2f} ms against a measured link floor of {b["e2e"]["link_floor_ms"]:.2f} ms for the same bytes.
Earlier in this round (commit 338cf7a): 0.2449 ms/step.  Round 1: 0.3144 ms/step, step roofline 0.205, dominant stage 0.28
with 450 MB of traffic (2.02x algorithmic).

| kernel | launch (us) | share | DRAM read / write (MB) | DRAM % | L2 hit % | L1 hit % | L2->SM (TB/s) | issue active % | warps active % | regs | stalls long_sb / barrier / short_sb |
|---|---|---|---|---|---|---|---|---|---|---|---|
{table}

Sum of `ls_*` kernels per step (ncu, serialised, cold caches): {tot:.1f} us.  In situ (CUPTI, graph replay, `r02_timeline.txt`):
span {span:.1f} us per step under the profiler, idle {idle:.1f} us; the dependent chain is camera {t("camera_transform"):.1f} -> index {t("index"):.1f} ->
scan {t("tile_totals") + t("tile_scan"):.1f} -> place {t("place"):.1f} -> canon {t("canon"):.1f} -> splat {t("splat_fwd"):.1f} -> gather {t("bwd_gather"):.1f} -> epilogue {t("bwd_epilogue"):.1f} =
{chain_us:.1f} us; histogram zeroing ({t("zero_counts"):.1f}) runs under the camera transform, softmax ({t("ls_softmax_kernel"):.1f}) and the NHWC staging of the
features ({t("to_nhwc"):.1f}) on side streams under index + scan.  Backward = gather + epilogue = {t("bwd_gather") + t("bwd_epilogue"):.1f} us
(earlier in the round: 69.8 + 17.0 with two epilogue kernels; round 1: 137 us with the transposer).

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
                ("r02_bench_overlap_bwd.json", "`LS_OVERLAP_BWD=1`: epilogue launched as the gather's programmatic dependent, per-image arrival counters"),
                ("r02_bench_gather_tma.json", "`LS_GATHER_TMA=1`: gradient rows through TMA gather4 into mbarrier-completed shared-memory stages instead of LDG.128"),
                ("r02_bench_staged_epilogue.json", "`LS_SOFTMAX_BWD_STAGED=1`: the two staged epilogue kernels (softmax backward || layout) instead of the thread-per-pixel one"),
                ("r02_bench_stress.json", "stress: B=32, 6 cams, D=96, 400x400 (configs[3]); round 1: 2.01 ms"),
                ("r02_bench_stress_bf16.json", "stress, bf16")):
    if not have(f):
        continue
    x = line(f)
    md += f'| `{f}` | {desc} | {x["ms_per_step"]:.4f} | {x["value"]:.0f} | {x["roofline"]["frac"]:.3f} / {x["roofline_step"]["frac"]:.3f} |\n'
sc = b["static_rig_cache"]
nc = b.get("nchw_compat")
md += f'''| `r02_bench.json` key `static_rig_cache` | opt-in static-rig cache (index structures reused, weights refreshed; bit-identical) | {sc["ms_per_step"]:.4f} | {sc["value"]:.0f} | - |
'''
if nc:
    md += f'''| `r02_bench.json` key `nchw_compat` | the NCHW path measured inside the default run (gradients bit-identical: {nc["gradients_bit_identical"]}) | {nc["ms_per_step"]:.4f} | {nc["value"]:.0f} | - |
'''
md += f'''| `r02_bench.json` key `gpu_reference` | the reference's torch op chain on the same GPU, full batch | {b["gpu_reference"]["ms_per_step"]:.0f} | {b["gpu_reference"]["value"]:.1f} | - |
| `r02_bench_reference.json` | `--impl reference`: the same on {ref["cpu_baseline"]["cores"]} host threads, FULL batch of 16 | {ref["ms_per_step"]:.0f} | {ref["value"]:.2f} | - |

'''
if all(have("r02_bench_%s.json" % x) for x in ("n2", "n4", "n8", "train", "train8", "train_reference")):
    n2, n4, n8 = [line("r02_bench_%s.json" % x) for x in ("n2", "n4", "n8")]
    t8, t1, tr = line("r02_bench_train8.json"), line("r02_bench_train.json"), line("r02_bench_train_reference.json")
    md += f'''Multi-GPU (torchrun, one rank per GPU, per-GPU batch 16 for the lift-splat and 12 for training; `r02_bench_n2/n4/n8.json`:
separate `gpurun --gpus N` calls at commit 146129d (before the 32-warp / L1-prefetch gather: the lift-splat step was 0.234 ms there; training, agent and
reference-arm lines: commit ed5b6c2 of this session, same harness); `r02_bench_train8.json` / `r02_nccl_n8.txt`: the 20-step training run at 8 GPUs,
taken earlier in the round at commit 129717f - the harness has not changed since and the lift-splat is 0.5 % of its step):

| GPUs | lift-splat samples/s (device) | ms/step | efficiency | e2e samples/s (ms; link floor) | train samples/s (ms/step) | train efficiency |
|---|---|---|---|---|---|---|
| 1 | {b["value"]:.0f} | {b["ms_per_step"]:.4f} | 1.00 | {b["e2e"]["value"]:.0f} ({b["e2e"]["ms_per_step"]:.2f}; {b["e2e"]["link_floor_ms"]:.2f}) | {b["train"]["samples_per_s"]:.1f} ({b["train"]["ms_per_step"]:.1f}) | 1.00 |
'''
    for n, x in ((2, n2), (4, n4), (8, n8)):
        md += (f'| {n} | {x["value"]:.0f} | {x["ms_per_step"]:.4f} | {x["value"] / n / b["value"]:.3f} | {x["e2e"]["value"]:.0f} ({x["e2e"]["ms_per_step"]:.2f}; {x["e2e"]["link_floor_ms"]:.2f}) | '
               f'{x["train"]["samples_per_s"]:.1f} ({x["train"]["ms_per_step"]:.1f}) | {x["train"]["samples_per_s"] / n / b["train"]["samples_per_s"]:.3f} |\n')
    md += f'''| 8 (`--workload train --steps 20`) | - | - | - | - | {t8["value"]:.1f} ({t8["ms_per_step"]:.1f}) | {t8["value"] / 8 / t1["value"]:.3f} |

* The lift-splat has no data-path collective: device-timed efficiency is 1.00 by construction (each rank's own step time).
* Training: DDP all-reduces 78.4 MB of gradients per step (19.6 M parameters) inside the timed region; the step grows from
  52.7 ms (1 GPU) to 58-60 ms (2-8 GPUs): efficiency 0.88 (12 timed steps inside the default line) to 0.92 (20-step run),
  above the 0.85 target.  NCCL reports no NVLS on these VMs (`r02_nccl_n8.txt`); what is lost is the exposed tail of the
  all-reduce after the last bucket (the camera-encoder trunk's gradients are produced last) plus per-rank jitter of the
  3 000+ small kernels of the stock PyTorch stack; the lift-splat library is 0.5 % of the step.
* With the reference's own lift-splat ops in the same stack the 1-GPU step is {tr["ms_per_step"]:.0f} ms ({tr["value"]:.1f} samples/s):
  the library makes the full training step {tr["ms_per_step"] / t1["ms_per_step"]:.1f}x faster.
* e2e does not scale past 2-4 GPUs on these hosts: {b["e2e"]["ms_per_step"]:.1f} / {n2["e2e"]["ms_per_step"]:.1f} / {n4["e2e"]["ms_per_step"]:.1f} / {n8["e2e"]["ms_per_step"]:.1f} ms per step at 1 / 2 / 4 / 8 ranks for 2 x 206 MB per
  rank.  The hosts are single-NUMA VMs (`numa_node` = -1 for every GPU, so the NUMA binding in `bench.py` is a no-op);
  `e2e.link_floor_ms` (the same bytes copied both ways by all ranks at once, no kernels: {b["e2e"]["link_floor_ms"]:.1f} / {n2["e2e"]["link_floor_ms"]:.1f} / {n4["e2e"]["link_floor_ms"]:.1f} / {n8["e2e"]["link_floor_ms"]:.1f} ms) is what PCIe
  plus the host memory system allow - the step is within 2-6 % of it at every rank count.

'''
if have("r02_bench_agent.json") and have("r02_bench_agent_reference.json"):
    ag, agr = line("r02_bench_agent.json")["latency"], line("r02_bench_agent_reference.json")["latency"]
    md += f'''Agent tick (B=1, `r02_bench_agent.json`, 1000 iterations): as one CUDA graph p50 {ag["graph"]["wall_ms"]["p50"]:.2f} ms / p99 {ag["graph"]["wall_ms"]["p99"]:.2f} ms
wall ({ag["graph"]["device_ms"]["p50"]:.2f} / {ag["graph"]["device_ms"]["p99"]:.2f} device); stream launches {ag["stream"]["wall_ms"]["p50"]:.1f} / {ag["stream"]["wall_ms"]["p99"]:.1f} ms; with the reference's torch
lift-splat ops in the same stack {agr["stream"]["wall_ms"]["p50"]:.1f} / {agr["stream"]["wall_ms"]["p99"]:.1f} ms (its host syncs prevent graph capture).  Paper: 74.92 ms on a Quadro RTX 5000.

'''
ga, ep, sp, ca = find(k, "ls_bwd_gather_occ"), find(k, "ls_bwd_epilogue"), find(k, "ls_splat_fwd_direct"), find(k, "ls_canon")
md += f'''## Reading

* **Backward = two kernels.**  The gather reads the channels-last gradient rows in place ({ga["dram_read_MB"]:.0f} MB of DRAM reads: 118 MB of
  distinct rows + feature rows + the pixel-major index, essentially every byte once; L1 hit {ga["l1_hit_pct"]:.0f} %, L2 hit {ga["l2_hit_pct"]:.0f} %,
  {ga["issue_active_pct"]:.0f} % issue-active, 32 warps/SM at {int(ga["regs"])} registers: each batch of eight dot products is reduced right away, which
  is what lets eight rows in flight per half-warp fit 64 registers; at 80 registers / 24 warps the kernel ran 2 us longer).  The epilogue - softmax backward and the NHWC -> NCHW fix-up of
  `grad_feat`, formerly two kernels staged through shared memory (17 us in situ, ~80 instructions per element) - is ONE
  thread-per-pixel kernel now ({ep["dram_read_MB"]:.0f} + {ep["dram_write_MB"]:.0f} MB, {ep["dram_pct"]:.0f} % of the DRAM peak, {ep["issue_active_pct"]:.0f} % issue-active, {int(ep["regs"])} registers): same bits, {t("bwd_epilogue"):.1f} us in situ.
  DRAM traffic of the whole backward stage: {bwd_traffic / 1e6:.1f} MB against 222.6 MB algorithmic ({bwd_traffic / 222.56e6:.2f}x; round 1: 450 MB, 2.02x).
* **What was tried on the backward and lost** (all built and timed, numbers in DESIGN.md section 5): the epilogue inside the gather
  (per-pixel softmax backward in the half-warp that owns the pixel + the layout fix-up by the last of every eight column-CTAs
  to finish: 118 us instead of 95, `__threadfence` invalidates L1 and the fix-up CTAs serialise their L2 round trips); the
  epilogue as the gather's programmatic dependent waiting on per-image arrival counters (`LS_OVERLAP_BWD=1`, kept, tested:
  hides 10 us of the epilogue and slows the gather's last wave by the same amount); a dead-point test per batch of 8 rows (+4 us).
* **Forward splat**: direct row stores, {sp["dram_read_MB"]:.1f} + {sp["dram_write_MB"]:.1f} MB, L1 hit {sp["l1_hit_pct"]:.1f} %, {int(sp["regs"])} registers, 640 B of shared memory; the cell
  flush is branch-free now (predicated stores, one multiply-add for the row address: 616 -> 520 SASS instructions), which did not move
  its time ({t("splat_fwd"):.1f} us in situ): the kernel sits at ~63 % of the L1 data-pipe wavefront peak.  Canonical ordering INSIDE the splat
  (shared-memory run per tile, no `recs_sorted` round trip, no separate kernel) was built and measured: splat 59 -> 79 us for the
  21 us kernel it removes - the ordering is bound by its per-tile chain of dependent loads wherever it runs; not kept.
  Forward stage traffic {fwd_traffic / 1e6:.0f} MB for 193 MB algorithmic.
* **The integer pipeline** (index {t("index"):.1f} + scan {t("tile_totals") + t("tile_scan"):.1f} + place {t("place"):.1f} + canon {t("canon"):.1f} = {t("index") + t("tile_totals") + t("tile_scan") + t("place") + t("canon"):.1f} us; zeroing hidden): placement lost a third
  of its instructions (block-uniform 64-bit bases, 32-bit offsets, unconditional clamped loads: 27 -> 24 us).  Canon is latency-bound
  per thread, not k^2-bound (10 compares per record on average): 64 / 128 / 256 / 512 / 1024 threads per tile give 73 / 36 / 21 / 30 /
  52 us; ordering a tile's cells by size so that a warp's compare loops have equal trip counts made it slower (27 us).  The
  opt-in static-rig cache replaces index, scan, place and canon by two streaming refresh kernels.
* **What bounds the two row gathers** (`r02_ablation.txt`: developer builds `LS_ABLATE=1/2` of the same kernels, timed in the same
  step): backward gather 73.8 us shipped / 69.0 us with all its memory traffic and a quarter of its arithmetic / 37.1 us with
  its arithmetic and no row traffic; canon + forward splat 80.0 / 80-83 / 58.6 us.  Both run in the time of their access pattern
  alone - one 256-byte row per record from L2, 636 MB per kernel.  `r02_tma_gather_bench.txt` (`tools/tma_gather_bench.cu`) is
  the chip's throughput for exactly that pattern: 2.49 M uniformly random rows of a 164 MB table at 6.9 / 8.1 / 9.2 TB/s with
  24 / 32 / 48 warps per SM of `LDG.128`, 8.2-8.6 TB/s through TMA gather4 (`UTMALDG.2D.GATHER4`), 14.5-16.5 TB/s from an
  L2-resident table.  The gather moves its 636 MB at 9 TB/s, the splat at 11 TB/s: 0.75-1.0 of what the hardware delivers for
  random row gathers; the HBM fraction is low because the algorithmic bytes are a third of that row traffic.
* **Measured and rejected in the last session** (A/B logs `r02_ab_*.txt`): canon ranking from a shared-memory copy of the tile's
  keys with warp-uniform 16-byte broadcast windows (+2.3 to +14 us); bulk L2 prefetch of the gradient 1-8 samples ahead
  (`UBLKPF.L2`: +-0.1 us - DRAM latency is not what the gather waits for); the gradient rows through TMA gather4 into
  mbarrier-completed shared-memory stages (`ls_bwd_gather_tma_kernel`, `LS_GATHER_TMA=1`: bit-identical, 91 vs 74 us; stays
  in the library as the opt-in Blackwell-native variant); L1 prefetch two batches ahead in the gather (73 vs 69.7 us) and the
  same one-window-ahead L1 prefetch in the forward splat (records two windows ahead, 79 registers: splat +10 us).
* **What did help in the last session**: the gather at 32 warps per SM (each batch of eight dot products reduced right away:
  64 registers, `r02_ab_gather_32warps.txt`: 74.0 -> 71.8 us) and `prefetch.global.L1` (SASS `CCTL.E.PF1`) of the next batch's
  rows by the lanes that hold their records (`r02_ab_gather_l1_prefetch.txt`: 71.7 -> 69.7 us): the batch in flight lives in
  registers, the one behind it in L1 - more bytes in flight per SM without registers.  Same bits in every variant (sha256 of
  BEV + gradients printed by `tools/ab_sha.sh`).
* SASS: `UBLKCP.G.S` (bulk async copy shared -> global, the TMA path) is in `ls_splat_fwd_kernel<.., LS_OUT_NHWC_BULK, 64>`
  (`cuobjdump -sass libls_b200.so | grep UBLKCP`: fp32 and bf16 variants); `griddepcontrol.launch_dependents` (ACQBULK / PDL trigger)
  in `ls_camera_transform_kernel` and, with `LS_OVERLAP_BWD=1`, in the gather; `UTMALDG.2D.GATHER4` + `SYNCS` (mbarrier) in
  `ls_bwd_gather_tma_kernel` (fp32 and bf16 features).
'''
open(P + "r02_summary.md", "w").write(md)
print("written", len(md))
