"""GPU, N = 2 (skipped on a single-GPU box): the staged prove across two ranks equals the single-GPU proof."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_prove_matches_single_gpu():
    from za_b200 import _lib
    if _lib.lib().za_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ZA_BENCH_PRINT_PROOF="1")
    one = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "3", "--log-m", "16", "--no-cpu-baseline", "--no-sub"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", "bench.py", "--gpus", "2", "--steps", "1", "--warmup", "3", "--log-m", "16", "--no-cpu-baseline"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stderr[-2000:]
    p1 = json.loads(one.stdout.strip().splitlines()[-1])["proof_hex"]
    p2 = json.loads(two.stdout.strip().splitlines()[-1])["proof_hex"]
    assert p1 == p2
