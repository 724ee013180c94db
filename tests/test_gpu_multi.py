"""2-GPU face exchange (peer-memory push, and the NCCL send/recv fallback) against the
single-domain oracle (skipped on 1-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", ["boxper3d", "rotated3d", "drude", "lorentz", "boxper3d-nccl"])
def test_two_gpu_parity(case):
    """3D periodic box, the same with randomly rotated element frames, and the 2D drude / lorentz tests (PML + incident field + ADE), cut over
    two ranks with the NCCL face exchange"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "drivers", "mgpu_parity.py"), case]
    env = dict(os.environ)
    if case.endswith("-nccl"):   # the fallback transport (default: stores into peer memory)
        cmd[-1] = case[:-5]
        env["NEKCEM_B200_P2P"] = "0"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
