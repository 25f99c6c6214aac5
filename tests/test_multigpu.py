"""N > 1 through the product's own collective (gtlc_comm_* + the ncclReduce inside gtlc_render)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_entry_points_load_nccl_without_a_gpu():
    """gtlc_comm_get_unique_id opens libnccl.so.2 lazily (no link-time dependency) and needs no device."""
    from gatling_b200 import gi
    uid = gi.comm_get_unique_id()
    assert len(uid) == gi.COMM_ID_BYTES and any(uid)
    assert gi._lib().gtlc_comm_world_size() == 0


@pytest.mark.gpu
def test_sharded_render_reduces_to_the_single_gpu_image():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "MULTIGPU OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
