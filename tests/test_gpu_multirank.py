"""GPU, >= 2 devices: the NCCL path of the library (sharded rows, all-reduce of [sums|counts|inertia], kmeans++
owner selection) must reproduce the single-rank fit.  Skipped on a one-GPU box; the protocol itself is covered on
CPU by tests/test_dist_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_fit_equals_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multirank_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MULTIRANK_RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1][len("MULTIRANK_RESULT "):])
    assert res["ok"], res
