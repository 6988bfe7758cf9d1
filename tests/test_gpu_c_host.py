"""GPU: the C ABI used from plain C (no Python, no torch in the client): tests/c/host_roundtrip.c
drives the host pipeline and checks itself against the C oracle."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_roundtrip(tmp_path):
    from oracle import msda_oracle
    from pctrans_b200 import build
    lib = build.build_library()
    ora = msda_oracle.build()
    exe = str(tmp_path / "host_roundtrip")
    libdir, oradir = os.path.dirname(lib), os.path.dirname(ora)
    cmd = ["/usr/bin/gcc", "-O1", "-std=c11", os.path.join(ROOT, "tests", "c", "host_roundtrip.c"),
           "-I", os.path.join(ROOT, "include"), "-L", libdir, "-L", oradir,
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}", "-Wl,-rpath,/usr/local/cuda/lib64",
           "-lpctrans_msda", "-lmsda_oracle", "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "C_HOST_ROUNDTRIP PASS" in proc.stdout, proc.stdout


def test_plain_c_client_of_the_linear_abi(tmp_path):
    """include/pct_linear.h from plain C with CUDA-runtime device buffers (tests/c/linear_client.c)."""
    from pctrans_b200 import build
    lib = build.build_library()
    exe = str(tmp_path / "linear_client")
    libdir = os.path.dirname(lib)
    cmd = ["/usr/bin/gcc", "-O1", "-std=c11", os.path.join(ROOT, "tests", "c", "linear_client.c"),
           "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-L", libdir,
           "-L", "/usr/local/cuda/lib64", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64",
           "-lpctrans_msda", "-lcudart", "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "C_LINEAR_CLIENT PASS" in proc.stdout, proc.stdout
