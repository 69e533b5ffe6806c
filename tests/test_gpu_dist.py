"""Sharded calibration on 2 GPUs of one node (NCCL + the in-kernel NVLink exchanges of csrc/peer.cuh): ranks must
hold bit-identical weights, layers 1-4 must equal the unsharded run of the same job, every layer must lie in the
reference's own ensemble range (tools/dist_check.py, which also runs the sharded alpha refinement).  Skipped on a
single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu` runs it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_gpu_calibration_matches_unsharded(peer):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    # a short NCCL watchdog: a mismatched collective must fail this test in a minute, not hold two GPUs for ten
    env = dict(os.environ, EFFQ_PEER=peer, MASTER_ADDR="127.0.0.1", TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC="120",
               EFFQ_DIST_TIMEOUT_S="90")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if peer == "1" else "29534", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=400)
    print(out.stdout[-6000:])
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        open(os.path.join(ROOT, "gpurun_out", f"r02_dist_check_2gpu_peer{peer}.log"), "w").write(out.stdout)
    assert out.returncode == 0
    assert "DIST OK world=2" in out.stdout and "ranks hold identical weights" in out.stdout
    assert "TUNE DIST OK world=2" in out.stdout
