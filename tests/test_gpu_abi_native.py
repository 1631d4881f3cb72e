"""GPU: a stand-alone C++ program (no Python, no PyTorch) drives libsyngular_b200.so through include/syngular_b200.h only."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_consumer_of_the_c_abi(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "syngular_b200")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi", "abi_smoke.cu"), "-L", libdir, "-lsyngular_b200", "-Xlinker", "-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI SMOKE OK" in r.stdout, r.stdout + r.stderr
