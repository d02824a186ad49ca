"""Builds and runs the C++ host-mirror test (tests/cpp/test_cpb200.cpp over include/cpb200.hpp) on the GPU and
checks the KAT it prints against tests/golden/reference_kats.json."""
import os
import subprocess

import pytest

from helpers import ROOT, kats

pytestmark = pytest.mark.gpu


def test_cpp_mirror():
    out_dir = os.path.join(ROOT, "tests", "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "test_cpb200")
    lib_dir = os.path.join(ROOT, "crypto_primitives_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_cpb200.cpp"),
                           "-L", lib_dir, "-l:libcpb200.so", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("CRH([0,1,2])")][0]
    got = int("".join(line.split("=")[1].split()), 16)
    assert got == int(kats()["sponge"]["squeeze3"][0])                 # R/sponge/poseidon/mod.rs:388-393
