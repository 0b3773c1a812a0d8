"""The XU-free fast paths (hoisted-reciprocal exact division, magic-number rounding, bit-trick binning) against the
plain IEEE formulation, brute force on the device: ~5e9 operand pairs including every tie and near-tie pattern."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fast_paths_equal_ieee_formulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this machine")
    exe = str(tmp_path / "fastdiv_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "fastdiv_check.cu")], check=True)
    res = subprocess.run([exe, "4096"], capture_output=True, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "bad_div=0 bad_round=0 bad_qdq=0 bad_bin=0" in res.stdout
