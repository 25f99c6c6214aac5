"""CPU, world_size 2, gloo: the N>1 host logic of the benchmark -- tile sharding ("tile % world == rank"),
zero-filled foreign pixels and ONE sum-reduce of the HDR accumulator to rank 0, and the alternative of
the same row, sample-range sharding (every pixel, samples [spp*rank/world, spp*(rank+1)/world)) -- exercised
with the CPU oracle as the renderer (the CUDA renderer is covered by test_shard_sum_equals_full and
test_sample_range_shards_sum_to_the_image on the GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gatling_b200 import scenes
from oracle import orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = scenes.cornell(48, 32, spp=2, max_bounces=3)
    part, _, rays = orc.OracleScene(d).render(shard=(rank, world, 8))
    t = torch.from_numpy(part.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    counts = torch.tensor([float(rays[0] + rays[1])], dtype=torch.float64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
        np.save(out_path + ".rays.npy", counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_shard_and_reduce(tmp_path):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = scenes.cornell(48, 32, spp=2, max_bounces=3)
    full, _, rays = orc.OracleScene(d).render()
    got = np.load(out)
    assert np.array_equal(got, full)                      # disjoint tiles: x + 0 is exact
    assert int(np.load(out + ".rays.npy")[0]) == rays[0] + rays[1]


def _sample_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = scenes.cornell(40, 28, spp=5, max_bounces=3)          # 5 samples over 2 ranks: [0, 2) and [2, 5)
    o = orc.OracleScene(d)
    shard = (rank, world, o.SHARD_SAMPLES)
    one, _, rays = o.render(shard=shard)
    two, _, _ = o.render(sample_offset=5, prev=one, shard=shard)          # progressive: every rank blends its own partial
    t = torch.from_numpy(np.stack([one, two]))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    counts = torch.tensor([float(rays[0] + rays[1])], dtype=torch.float64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
        np.save(out_path + ".rays.npy", counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_range_shard_and_reduce(tmp_path):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_sample_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = scenes.cornell(40, 28, spp=5, max_bounces=3)
    o = orc.OracleScene(d)
    one, _, rays = o.render()
    two, _, _ = o.render(sample_offset=5, prev=one)
    got = np.load(out)
    assert np.all(got[..., 3] == 1.0)                                    # w: 1 on rank 0, 0 elsewhere
    for g, full in zip(got, (one, two)):                                  # same samples, summed in another order
        assert full[..., :3].mean() > 0.05
        assert np.allclose(g[..., :3], full[..., :3], rtol=2e-6, atol=1e-7)
    assert not np.array_equal(one, two)
    assert int(np.load(out + ".rays.npy")[0]) == rays[0] + rays[1]        # the ranks trace exactly the single-process rays
