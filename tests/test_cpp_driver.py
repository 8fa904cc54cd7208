"""tests/cpp/abi_parity.cpp: a C++ host (no Python between it and the libraries) drives the C ABI the way FullSystem would
-- makeImages, PixelSelector::makeMaps, ImmaturePoint construction, traceNewCoarse over two frames, FullSystem::optimize(6)
-- on two libraries and compares them call by call.  CPU: the driver builds against include/sosba.h and the oracle agrees
with itself (the case is a real one: > 150 points, > 300 active residuals).  GPU: libsosba.so against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "abi_parity.cpp")
ORC = os.path.join(ROOT, "oracle", "_build", "liborc_parity.so")


@pytest.fixture(scope="module")
def driver(built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "abi_parity")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe, SRC, "-ldl"], check=True, cwd=ROOT)
    return exe


def _run(exe, lib_a, pre_a):
    p = subprocess.run([exe, lib_a, pre_a, ORC, "orc"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ABI_PARITY OK" in p.stdout, p.stdout + p.stderr
    return p.stdout


def test_cpp_driver_oracle(driver):
    out = _run(driver, ORC, "orc")
    assert "DIFFERENT" not in out


@pytest.mark.gpu
def test_cpp_driver_cuda_vs_oracle(driver, built):
    out = _run(driver, built.LIB_PATH, "sosba")
    assert "DIFFERENT" not in out      # pyramid, selection, immature points and both traces are bit-exact
