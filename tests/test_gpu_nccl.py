"""Two-GPU run of tests/nccl_band_flux.py (the multi-GPU row of SURVEY.md 8e with the real CUDA path and NCCL);
skipped on boxes with a single GPU, where tests/test_multirank_gloo.py covers the host logic."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_band_flux_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "nccl_band_flux.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL_BAND_FLUX_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_multi_gpu_band_flux_from_cpp_over_the_c_abi():
    """tests/cpp/nccl_band_cabi.cpp: a C++ host (no torch) shards spectral points over every GPU of the box, solves
    with the fused band integration and finishes with b200disort_allreduce_band on ncclCommInitAll communicators."""
    exe = os.path.join(HERE, "cpp", "build", "nccl_band_cabi")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run build() first")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL_BAND_CABI_OK" in r.stdout
