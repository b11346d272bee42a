"""world_size-2 gloo test of the data-parallel plumbing: samples are
independent, so every rank computing its own slice and gathering gives exactly the
unsharded result.  The per-rank compute here is the CPU port (no GPU in this suite); on
GPUs bench.py runs the same sharding over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import frustum_of, grid_of
from e2e_parking_carla_b200 import sharding
from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_encoder_outputs, make_rig, make_upstream_grads


def test_shard_bounds_cover_the_batch():
    for batch in (1, 2, 3, 7, 16, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import torch_port as tp
    shape = LiftSplatShape(batch=3, channels=4)            # 3 samples over 2 ranks: uneven split
    intr, extr = make_rig(3, 4, jitter=True, seed=5)
    feat, logits = make_encoder_outputs(shape, seed=4)
    gb, gp = make_upstream_grads(shape, seed=4)
    res, start, dim = grid_of(shape)
    a = (torch.from_numpy(frustum_of(shape)), torch.from_numpy(start), torch.from_numpy(res), torch.from_numpy(dim))
    sl = lambda t, pc: sharding.shard_batch(t, shape.cams, world, rank, per_camera=pc)
    bev, prob, gf, gl = tp.fwd_bwd_step(sl(feat, True), sl(logits, True), sl(intr, False), sl(extr, False), *a,
                                        sl(gb, False), sl(gp, True))
    full_bev = sharding.gather_batch(bev, 3, shape.cams, per_camera=False)
    full_gf = sharding.gather_batch(gf, 3, shape.cams, per_camera=True)
    slowest = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        ref = tp.fwd_bwd_step(feat, logits, intr, extr, *a, gb, gp)
        np.save(os.path.join(out_dir, "ok.npy"),
                np.array([torch.equal(full_bev, ref[0]), torch.equal(full_gf, ref[2]), slowest == 10.0 + world - 1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all(), ok


def _ddp_worker(rank, world, port, out_dir):
    """One DDP step of the training harness (harness/parking_stack.py) on CPU over gloo, with the
    reference's torch lift-splat ops standing in for the CUDA library: the plumbing bench.py
    --workload train runs over NCCL (BASELINE.json configs[2])."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from harness.parking_stack import Losses, ParkingStack, count_parameters, default_cfg, synthetic_batch
    cfg = default_cfg("cpu")
    cfg.final_dim = [64, 64]                    # small images: the stack is resolution-agnostic up to the BEV
    torch.manual_seed(42)
    model = ParkingStack(cfg, lift_splat="torch")
    net = DDP(model, gradient_as_bucket_view=True, static_graph=True)
    crit = Losses(cfg, native=False)
    data = synthetic_batch(cfg, 1, "cpu", seed=rank)
    loss = crit(net(data), data)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    unused = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    losses = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1))
    if rank == 0:
        np.save(os.path.join(out_dir, "ddp.npy"),
                np.array([float(torch.equal(gathered[0], gathered[1])), float(torch.isfinite(flat).all()),
                          float(len(unused) == 0), float(losses[0] != losses[1]), float(count_parameters(model))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ddp_training_step(tmp_path):
    """Gradients are averaged across ranks (identical after backward although the ranks saw
    different batches), every trainable parameter receives one (no unused parameters: DDP needs no
    find_unused_parameters, unlike the reference whose BevEncoder.layer4 is never called), and the
    stand-in has the reference's live parameter count (28.0 M - 8.39 M of layer4 = 19.6 M)."""
    world = 2
    mp.spawn(_ddp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    same, finite, all_used, different_batches, params = np.load(tmp_path / "ddp.npy")
    assert same == 1 and finite == 1 and all_used == 1 and different_batches == 1
    assert 19.0e6 < params < 20.2e6
