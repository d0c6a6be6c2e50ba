"""The C++ mirror of the reference API (include/vrfs_b200.hpp) - the compiled-language host side that stands in for the Rust
shim (no Rust toolchain in the image): builds on any box, runs the reference's round-trip tests + an upstream vector on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ark_ec_vrfs_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_api_mirror")


def build():
    from ark_ec_vrfs_b200 import build as b
    b.build()
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_api_mirror.cpp"),
                    "-L" + PKG, "-lvrfs_b200", "-Wl,-rpath," + PKG], check=True)


def test_cpp_mirror_builds_against_the_c_abi():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_roundtrips_and_golden_vector():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all ok" in r.stdout, r.stdout + r.stderr
