"""Multi-GPU checks (skipped on a single-GPU box): launch tests/dist_gpu_check.py and bench.py
under torchrun with 2 ranks."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, port=29517):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_row_block_sharded_ice_two_gpus():
    r = _torchrun(2, os.path.join("tests", "dist_gpu_check.py"))
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_bench_two_gpus_chromosome_sharding():
    r = _torchrun(2, "bench.py", "--gpus", "2", "--pairs", "20000000", "--steps", "2", "--warmup", "3", port=29519)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["roofline"]["frac"] > 0
