"""GPU test: the C++ host mirror of kyber.Group/Point/Scalar + pairing.Suite (kyber_b200/host/kyber_b200.hpp)
passes a conformance run modelled on the reference's testGroup / pairing / sign/bls tests (tests/cpp/test_suite.cpp)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_conformance():
    exe = os.path.join(ROOT, "tests", "cpp", "test_suite")
    if not os.path.exists(exe):
        subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_suite.cpp"), "-o", exe,
                        "-L" + os.path.join(ROOT, "kyber_b200"), "-lb2kyber", "-Wl,-rpath," + os.path.join(ROOT, "kyber_b200"),
                        "-L/usr/local/cuda/lib64", "-lcudart"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout and r.stdout.count("ok ") >= 45
