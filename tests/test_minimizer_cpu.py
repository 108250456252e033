"""The minimizer helpers of alga_b200/csrc/common.cuh (experimental seed-index variant) are plain integer code:
tests/minimizer_check.cu compiles them for the host with nvcc and compares the sliding minimum with the from-scratch one."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not installed")
def test_sliding_minimizer_equals_window_minimizer(tmp_path):
    exe = str(tmp_path / "minimizer_check")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-o", exe, os.path.join(HERE, "minimizer_check.cu")],
                   check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout
