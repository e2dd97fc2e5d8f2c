"""The one collective of the path on real GPUs: the fused push all-gather (the step kernel's TMA stores write every
observation chunk into all ranks' gathered tensors over NVLink) must return exactly what the NCCL all-gather returns -- also
when one rank's consumer is slow and nothing but the protocol's own barrier keeps the ranks apart (two-slot ring,
include/rware_b200.h rw_step_gather).  Spawns 2 ranks with torchrun when at least 2 GPUs are visible."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_fused_push_all_gather_equals_nccl_even_with_a_slow_consumer():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MGPU_ENVS="8192", MGPU_SKEW_ENVS="2048", MGPU_SKEW_STEPS="200")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py")],
                         capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["fused_equals_nccl"] and rec["fused_equals_nccl_with_slow_consumer"], rec
