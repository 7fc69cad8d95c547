"""The C++ host mirror (include/psim_b200.hpp): compiles against the C ABI on CPU; on the GPU box the
reference's own quadtree tests, rewritten against it, run and pass."""
import os
import subprocess

import pytest

from helpers import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_kats.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_reference_kats")


def compile_mirror():
    from particlesim_b200 import _lib
    _lib.load()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.join(ROOT, "particlesim_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", BIN,
           "-L", libdir, "-lpsim_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return BIN


def test_cpp_mirror_compiles_and_links():
    compile_mirror()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_reference_tests_through_cpp_mirror(cuda_device):
    exe = compile_mirror()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all passed" in res.stdout
