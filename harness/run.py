"""Timed loops for BASELINE.json configs[2] (full training step under DDP) and configs[4]
(closed-loop agent inference latency), used by bench.py --workload train / agent."""
from __future__ import annotations

import os
import statistics
import time

import torch

from .agent_target import prev_target_point
from .parking_stack import Losses, ParkingStack, count_parameters, default_cfg, synthetic_batch


def _dist(world, device):
    if world <= 1:
        return None
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    return dist


def train_benchmark(per_gpu_batch, steps, warmup, rank, world, device, lift_splat="b200", channels_last=True,
                    e2e_steps=None):
    """One rank of the DDP training benchmark (trainer/pl_trainer.py:55-83,116-121 restated as a
    plain torch loop: forward, three losses, backward with NCCL gradient all-reduce overlapped by
    DDP's bucketing, Adam step).  Returns a dict (identical on all ranks for the timing fields)."""
    dist = _dist(world, device)
    cfg = default_cfg(device)
    torch.manual_seed(42)                                    # pl_train.py: seed_everything(42)
    model = ParkingStack(cfg, lift_splat=lift_splat, channels_last=channels_last).to(device)
    crit = Losses(cfg, native=lift_splat == "b200").to(device)
    net = model
    if dist is not None:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[device.index], gradient_as_bucket_view=True, static_graph=True)
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, weight_decay=cfg.weight_decay)
    data = synthetic_batch(cfg, per_gpu_batch, device, seed=rank)
    host = {k: v.cpu().pin_memory() for k, v in data.items()}

    def step(batch):
        opt.zero_grad(set_to_none=True)
        loss = crit(net(batch), batch)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(data)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(data)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    # end to end: the batch comes from pinned host memory every step, the loss goes back to the host
    n_e2e = e2e_steps or max(3, min(steps, 10))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(n_e2e):
        batch = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        last = float(step(batch).item())
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / n_e2e
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    params = count_parameters(model)
    return {"ms_per_step": ms, "samples_per_s": per_gpu_batch * world / (ms * 1e-3),
            "e2e_ms_per_step": ms_e2e, "e2e_samples_per_s": per_gpu_batch * world / (ms_e2e * 1e-3),
            "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": 4,
            "loss": last, "trainable_params": params, "allreduce_bytes_per_step": params * 4 if world > 1 else 0,
            "per_gpu_batch": per_gpu_batch, "lift_splat": lift_splat, "channels_last": channels_last}


def agent_benchmark(iters, warmup, device, lift_splat="b200", use_graph=True):
    """Closed-loop agent step without the CARLA server (agent/parking_agent.py:379-391): model.predict
    on one synthetic frame set (B=1, 4 cameras) + the next-target point of save_prev_target (:290-318)
    computed on the device (harness/agent_target.py) instead of a python 200x200 loop and fed back as the
    next tick's target point.  Per-iteration latency by CUDA events and by
    wall clock (the reference's own time.time() bracket, which includes the final device->host read)."""
    cfg = default_cfg(device)
    torch.manual_seed(42)
    model = ParkingStack(cfg, lift_splat=lift_splat).to(device).eval()
    data = synthetic_batch(cfg, 1, device, seed=0)
    data["gt_control"] = data["gt_control"][:, :1]            # BOS only (agent/parking_agent.py:470)
    x_res, y_res = cfg.bev_x_bound[2], cfg.bev_y_bound[2]

    def tick():
        tokens, seg, _, _ = model.predict(data)
        # save_prev_target (:290-318) on the device, bit-equal to the reference's python loop
        # (tests/test_agent_target.py); the next tick aims at it (:474-475), which closes the loop
        # inside the captured graph: the copy below feeds the graph's own static input
        target, _ = prev_target_point(seg, x_res, y_res, data["target_point"][0, :2])
        data["target_point"][0, :2].copy_(target)
        return tokens, target

    def measure(run_once):
        dev_ms, wall_ms = [], []
        for i in range(iters + warmup):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            tokens, centroid = run_once()
            e1.record()
            host = (tokens.cpu(), centroid.cpu())         # what the agent consumes next (syncs)
            t1 = time.perf_counter()
            if i >= warmup:
                dev_ms.append(e0.elapsed_time(e1))
                wall_ms.append((t1 - t0) * 1e3)
        dev_ms.sort()
        wall_ms.sort()
        q = lambda v, p: v[min(len(v) - 1, int(p * len(v)))]
        return {"device_ms": {"p50": q(dev_ms, 0.5), "p99": q(dev_ms, 0.99), "mean": statistics.fmean(dev_ms)},
                "wall_ms": {"p50": q(wall_ms, 0.5), "p99": q(wall_ms, 0.99), "mean": statistics.fmean(wall_ms)}}

    res = {}
    with torch.no_grad():
        for _ in range(warmup):
            tick()
        torch.cuda.synchronize()
        res["stream"] = measure(tick)
        if use_graph:
            # the whole tick as one CUDA graph (the library is capture-safe: no allocation, no host sync;
            # the target-pixel noise draws from the graph-registered generator)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = tick()
                torch.cuda.synchronize()

                def replay():
                    g.replay()
                    return out

                res["graph"] = measure(replay)
            except Exception as exc:
                res["graph_error"] = repr(exc)[:400]
    res["iters"], res["warmup"], res["lift_splat"] = iters, warmup, lift_splat
    return res
