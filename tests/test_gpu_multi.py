"""The multi-GPU path on real GPUs: one workload split by sharding.estimate_regions_sharded over two
ranks, one GPU each, host-side gather.  Needs two visible GPUs (`gpurun --gpus 2`); with one GPU the test says so."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_on_two_gpus_equals_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the world-size-2 logic is covered on CPU by tests/test_sharding.py)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617",
                          os.path.join(ROOT, "tests", "helpers", "sharded_ranks.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["equal"] and min(res["pieces_per_rank"]) >= 1 and res["reads_with_round3"] > 600, res
