"""Compiles and runs the C++ host-side mirror of the reference interface
(cqs_b200/host/b200_index.hpp) against libcqs_b200.so on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(out):
    lib_dir = os.path.join(ROOT, "cqs_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
                           "-o", out, "-L" + lib_dir, "-l:libcqs_b200.so", "-Wl,-rpath," + lib_dir])


def test_host_mirror_compiles(tmp_path):
    _compile(str(tmp_path / "host_mirror"))


@pytest.mark.gpu
def test_host_mirror_conformance(tmp_path):
    exe = str(tmp_path / "host_mirror")
    _compile(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host mirror OK" in r.stdout, r.stdout + r.stderr
