"""The C++ host mirror compiles against include/shodh_gpu.h (CPU check) and passes the reference-style
unit tests on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "_host_mirror_test")


def _build(sg):
    libdir = os.path.dirname(sg.capi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-o", EXE, SRC, f"-L{libdir}", "-lshodh_gpu", f"-Wl,-rpath,{libdir}",
           "-L/usr/local/cuda/lib64", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_host_mirror_compiles(sg):
    _build(sg)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_runs_reference_style_tests(sg):
    _build(sg)
    res = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "HOST_MIRROR_OK" in res.stdout
