"""CPU execution of the register-FFT / phase-correlation pass code
(csrc/fft_reg.cuh, csrc/fft_pass.cuh): the per-thread pieces are __host__
__device__, and tests/csrc/*_emul.cu run them thread by thread against float64
DFTs.  Needs nvcc (host compilation only), no GPU."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.parametrize("name", ["fft_emul", "pass_emul"])
def test_emulation(name, tmp_path):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / name)
    src = os.path.join(ROOT, "tests", "csrc", name + ".cu")
    inc = os.path.join(ROOT, "multiview_stitcher_b200", "csrc")
    subprocess.run([NVCC, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-I", inc, src, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.strip().endswith("OK")
