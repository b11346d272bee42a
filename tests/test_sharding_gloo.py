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
