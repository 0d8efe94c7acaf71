"""The C++ host mirror of the reference interface (stark-rs_b200/host/stark.hpp): builds on CPU, runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror")


def _build():
    import oracle as O
    import stark_rs_b200 as S
    O.build()
    S.build_library()
    cmd = ["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", EXE,
           "-L" + os.path.join(ROOT, "stark-rs_b200"), "-lstark_b200", "-L" + os.path.join(ROOT, "oracle"), "-l:liboracle.so",
           "-Wl,-rpath," + os.path.join(ROOT, "stark-rs_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def test_host_mirror_builds():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_runs():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout
