"""GPU (-m gpu): sharded state on >= 2 GPUs of one box (skipped on a 1-GPU box).  Launches
tests/dist_gpu_check.py under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_state_two_gpus(lib_built):
    import cunqa_b200 as cb
    ngpu = cb.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert "DIST PARITY OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
