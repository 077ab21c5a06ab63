"""GPU: multi-GPU particle sharding -- the sharded filter must reproduce the unsharded one.
Needs >= 2 GPUs for the real exchange (run with `gpurun --gpus 2`); with one GPU the world-size-1 group
still exercises the shard code path (combine, shard-aware scan, push kernel, finish)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(nproc, n_local, port, exchange="p2p", mode="philox"):
    env = dict(os.environ, SHARD_N_LOCAL=str(n_local), SHARD_EXCHANGE=exchange, SHARD_MODE=mode)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "shard_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "[shard_worker] OK" in p.stdout
    return p.stdout


def test_shard_world1():
    run_workers(1, 1 << 16, 29611)


@pytest.mark.parametrize("n_local", [1 << 16, 1 << 22])
def test_shard_world1_push_kernel_vs_oracle(n_local):
    """k_step_push + shard-aware scan + peer exchange (world 1) against the CPU oracle with supplied noise."""
    run_workers(1, n_local, 29621, mode="noise")


def test_shard_world2_push_kernel_vs_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    run_workers(2, 1 << 16, 29622, mode="noise")
    run_workers(2, 1 << 23, 29623, mode="noise")  # one filter of 2^24 particles over two GPUs


def test_shard_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = run_workers(2, 1 << 16, 29612)
    assert "cross_shard_fraction" in out
    run_workers(2, (1 << 20) + 2048 * 3, 29613)
    run_workers(2, 1 << 16, 29615, exchange="nccl")  # host-issued NCCL collectives on the filter stream


def test_shard_world_all_gpus():
    """Every GPU of the box (4 or 8): CUDA-IPC peer tables, ranges and the P2P push with more than two ranks."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 3:
        pytest.skip("needs >= 3 GPUs (gpurun --gpus 4|8)")
    out = run_workers(ng, 1 << 16, 29614)
    assert "cross_shard_fraction" in out
