"""GPU, more than one device: the sharded hot path over REAL NCCL (psim_comm_init + psim_step_sharded, one process per
GPU under torch.distributed.run; strict node centres, polar pass, locally-essential-tree exchange) must leave the
single-GPU state, bit for bit.  Skipped on a one-GPU box (NCCL refuses
two ranks on one device); the one-GPU protocol tests are tests/test_gpu_shard.py."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


CASES = [
    # orchestration, bodies, generator, theta, extra environment
    ("library", 200001, "electrolyte", 1.0, {"PSIM_LET_POISON": "1"}),
    ("python", 200001, "electrolyte", 1.0, {}),
    ("library", 300000, "clustered", 0.5, {"PSIM_LET_POISON": "1"}),      # deep unbalanced tree, wide opening
    ("library", 150001, "uniform_pm1", 0.7, {"PSIM_LET_POISON": "1"}),    # every body charged
    ("library", 200001, "electrolyte", 1.0, {"PSIM_LET_CAP": "100"}),     # send areas overflow: full all-gather fall-back
]


@pytest.mark.parametrize("how,n,gen,theta,extra", CASES)
def test_sharded_steps_over_nccl_equal_single_gpu(cuda_device, how, n, gen, theta, extra):
    g = _gpus()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if g >= 8 else (4 if g >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "check_sharded.py"),
           str(n), how, gen, str(theta)]
    # PSIM_LET_POISON: the records the locally-essential-tree exchange does not deliver are overwritten with NaN centres
    # and dangling pointers, so a walk that reached one could not give the single-GPU bits
    env = dict(os.environ, **extra)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(res.stdout[-3000:], res.stderr[-2000:])
    assert res.returncode == 0 and "== SINGLE: True" in res.stdout
