"""The C++ host mirror of the reference operator API (include/phare_b200/phare_b200.hpp): compiles everywhere;
on a GPU box the test program runs IonUpdater / Faraday / Ampere / Ohm through it and checks the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_operators")
EXE_AMR = os.path.join(ROOT, "build", "test_amr_ops")


def build(src="test_operators.cpp", exe=EXE):
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    cmd = ["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp", src),
           "-o", exe, "-L" + os.path.join(ROOT, "phare_b200/lib"), "-lphare_b200", "-L" + os.path.join(ROOT, "oracle"),
           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "phare_b200/lib"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_mirror_compiles_and_links():
    build()
    r = subprocess.run([EXE, "--compile-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "compiled" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_cpp_amr_mirror_compiles_and_links():
    """include/phare_b200/amr.hpp: the refine / coarsen policy classes of the reference over the C ABI"""
    build("test_amr_ops.cpp", EXE_AMR)
    r = subprocess.run([EXE_AMR, "--compile-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "compiled" in r.stdout


@pytest.mark.gpu
def test_cpp_amr_mirror_matches_oracle_on_gpu():
    build("test_amr_ops.cpp", EXE_AMR)
    r = subprocess.run([EXE_AMR], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
