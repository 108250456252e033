"""Sharded build on real peer memory (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed_keys", ["0", "1"])  # the peers derive the seeds of all reads / pull 12-byte seed records
@pytest.mark.parametrize("world", [2])
def test_sharded_build_matches_oracle(gpu, world, seed_keys):
    from alga_b200 import _lib

    if _lib.load().alga_gpu_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_worker.py"), "0.02"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, ALGA_SHARD_SEED_KEYS=seed_keys))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "match=True" in r.stdout
