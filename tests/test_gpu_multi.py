"""Multi-GPU paths on real devices; every test skips on a box with fewer than two GPUs.
  * one process, one context over several devices (nd.init([0,1])): shards, replicated broadcast operands, the row-panel split
    of a single large matmul and the global cholesky failure index — bit-for-bit against the one-device context;
  * one process per GPU with the results gathered by NCCL (the C5 configuration of BASELINE.json).
On a one-GPU box the world_size-2 logic is covered by the gloo test in test_host_logic.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_multi_device_context_matches_one_device_bit_for_bit():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multidev_check.py"), "2"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_devices"] == 2
    assert line["all_equal"], line


def test_sharded_svd_with_nccl_gather_matches_unsharded():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571",
                          os.path.join(ROOT, "tools", "sharded_svd_gather.py"), "1026"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["sharded_equals_unsharded_bits"] and line["max_rel_residual"] <= 1e-12 and line["max_orth_error"] <= 1e-12
