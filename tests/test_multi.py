"""Multi-rank path.  CPU: world_size-2 gloo run of the sharded algorithm (host logic + partition helpers of the C ABI).
GPU: 2 (or more) B200s through the library's NCCL exchange, skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, case, world, partition="region"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "multi_worker.py"), mode, case, partition]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_OK" in p.stdout, p.stdout[-2000:]
    return p.stdout


@pytest.mark.parametrize("case,world,partition", [("synth_small", 2, "walk"), ("synth_repeats", 2, "walk"), ("synth_dirty", 3, "walk"),
                                                  ("synth_small", 2, "region"), ("synth_repeats", 3, "region"), ("synth_dirty", 2, "region")])
def test_sharded_algorithm_cpu_gloo(case, world, partition):
    launch("cpu", case, world, partition)


@pytest.mark.gpu
@pytest.mark.parametrize("case,partition", [(c, "region") for c in ["toy_k3_w2", "synth_small", "synth_dirty", "synth_repeats", "mhc4", "synth:77:400000:11:3.0"]]
                         + [(c, "walk") for c in ["synth_small", "synth_dirty", "synth:77:400000:11:3.0"]]
                         + [("synth_small+d1", "region"), ("synth_repeats+d1", "walk")])
def test_multi_gpu_matches_oracle(case, partition):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    launch("gpu", case, min(n, 4) if case != "toy_k3_w2" else 2, partition)
