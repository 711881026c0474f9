"""Host-side check of the counter-based dropout hash (csrc/common.cuh): the per-run fast path the kernels use
(affine base + j * C1, tasr_hash_finish) must equal the generic pair hash for every pair of a run, because forward
and backward kernels of one dropout site may use either form (GEMM epilogue vs. GroupNorm-backward cast); the keep
rate must match 1 - p.  Compiled with nvcc, run on the CPU (host functions only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropout_hash_fast_path_matches_generic(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "dropout_hash_check")
    cmd = [nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-I", os.path.join(ROOT, "turkish_asr_model_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "tests", "host", "dropout_hash_check.cu")]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=60).stdout.split()
    bad, n, keep = int(out[0]), int(out[1]), float(out[2])
    assert n > 100000
    assert bad == 0
    assert abs(keep - 0.9) < 0.005
