#!/usr/bin/env python
"""Multi-GPU check of the in-library NCCL reduce, run under torchrun with one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multigpu_check.py

Every rank renders its tiles of the same scene; gtlc_render ends with ncclReduce(sum) to rank 0 before the device->host
copy.  Rank 0 then renders the whole frame alone and demands the two host images be BIT-IDENTICAL (x + 0 = x), for a
single call and for two progressive calls; then the same with sample-range sharding (every pixel, a share of the samples),
where the sum runs in another order and the images must agree to rounding.  torch.distributed (gloo) only carries the
128-byte NCCL id.
"""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from gatling_b200 import gi, scenes  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
dist.init_process_group("gloo")
assert gi.giInitialize(local_rank) == gi.GiStatus.Ok, gi.giLastError()
box = [gi.comm_get_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
gi.comm_init(box[0], rank, world)

d = scenes.cornell(200, 136, spp=6, max_bounces=5)
d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
d.rect_lights.append(scenes.RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
b = gi.BuiltScene(d)
ok = True
for tile in (16, 64):
    gi.set_shard(b.scene, rank, world, tile)
    one = b.render().copy()             # samples [0, 6)
    two = b.render().copy()             # progressive: samples [6, 12) blended in
    dist.barrier()
    if rank == 0:
        gi.set_shard(b.scene, 0, 1, tile)
        ref_one = b.render().copy()
        ref_two = b.render().copy()
        same = np.array_equal(one, ref_one) and np.array_equal(two, ref_two)
        lit = float(ref_two[..., :3].mean())
        print("tile %d: sharded over %d GPUs == single GPU: %s (mean %.4f)" % (tile, world, same, lit), flush=True)
        ok = ok and same and lit > 0.05 and not np.array_equal(ref_one, ref_two)
    dist.barrier()
gi.set_shard(b.scene, rank, world, gi.SHARD_SAMPLES)
one = b.render().copy()
two = b.render().copy()
dist.barrier()
if rank == 0:
    gi.set_shard(b.scene, 0, 1, 64)
    ref_one = b.render().copy()
    ref_two = b.render().copy()
    close = all(np.allclose(x[..., :3], y[..., :3], rtol=2e-6, atol=1e-7) and np.all(x[..., 3] == 1.0) for x, y in ((one, ref_one), (two, ref_two)))
    print("sample ranges: sharded over %d GPUs ~= single GPU: %s (max abs diff %.3g)" % (world, close, float(np.abs(two - ref_two).max())), flush=True)
    ok = ok and close
dist.barrier()
b.close()
gi.comm_destroy()
gi.giTerminate()
if rank == 0:
    print("MULTIGPU OK" if ok else "MULTIGPU MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
