"""Multi-GPU bit-identity on hardware: needs >= 2 B200s in one box (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu_hw.py -m gpu`).  The CPU-side logic of the same path is
covered by tests/test_multi_gpu_cpu.py (gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_scene_job_is_bit_identical_across_world_sizes(gpu, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py"), "2312"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout
