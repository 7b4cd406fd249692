"""Builds tests/cpp/test_services.cpp against include/sa_services.hpp + libsa_engine.so with g++ and runs
it: the compiled-language host layer (the stand-in for the reference's Java services) on a real GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "test_services")
    lib_dir = os.path.join(ROOT, "spectral_analyzer_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_services.cpp"), "-o", exe,
                    "-L", lib_dir, "-lsa_engine", "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_cpp_host_compiles_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_host_runs(tmp_path):
    out = subprocess.run([_build(tmp_path)], check=True, capture_output=True, text=True).stdout
    assert "cpp services ok" in out
