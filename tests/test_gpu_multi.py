"""§8(e): point shards over several GPUs (NCCL) reproduce the single-GPU optimize.  Needs >= 2 visible GPUs; the
host-side sharding logic is covered on CPU by test_cpu_oracle.py::test_point_shards_allreduce_gloo."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["small", "configB"])
def test_point_shards_nccl(built, which):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), which], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout
