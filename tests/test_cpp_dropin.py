"""The C++ drop-in front end (include/binary/algorithm/interval_tree.hpp): the reference's doctest cases
re-expressed in tests/cpp/test_dropin.cpp. CPU: it compiles and the host-only cases pass; GPU: all."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")


@pytest.fixture(scope="module")
def dropin_binary():
    r = subprocess.run(["make", "-C", CPP, "test_dropin"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(CPP, "test_dropin")


def test_front_end_compiles_and_host_side_cases_pass(dropin_binary):
    r = subprocess.run([dropin_binary], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_unit_tests_through_the_front_end_on_gpu(dropin_binary):
    r = subprocess.run([dropin_binary, "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
