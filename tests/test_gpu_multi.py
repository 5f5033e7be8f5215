"""Multi-GPU path on real devices: the batch sharded over one process per GPU, results gathered with NCCL.
Needs two visible GPUs; on a one-GPU box the world_size-2 logic is covered by the gloo test in test_host_logic.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_svd_with_nccl_gather_matches_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571",
                          os.path.join(ROOT, "tools", "sharded_svd_gather.py"), "1026"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["sharded_equals_unsharded_bits"] and line["max_rel_residual"] <= 1e-12 and line["max_orth_error"] <= 1e-12
