"""CPU check of the split Philox block the strip loop uses (mcx_common.cuh: philox_head / philox_tail): built with nvcc as a
host program -- no GPU involved -- it must equal the plain Philox4x32-10 block on random and edge inputs, and the plain block
must reproduce the Random123 known-answer vectors."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_philox_block_equals_the_plain_block(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "philox_split_check")
    src = os.path.join(ROOT, "tests", "aux", "philox_split_check.cu")
    build = subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "mismatches=0 kat0=1 kat1=1" in run.stdout
