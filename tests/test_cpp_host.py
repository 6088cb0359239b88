"""The C++ host mirror of reth's interface (reth_b200/host/reth_b200.hpp) exercised by tests/cpp/host_test.cpp:
reth's account_and_storage_trie / from_bundle_state / extension-node tests restated in C++ over the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_test")


def _ensure_built():
    if not os.path.exists(BIN):
        import __graft_entry__
        __graft_entry__.build()


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu():
    _ensure_built()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_cpp_host_fails_loudly_without_gpu():
    """No CPU fallback: without a device the Engine constructor throws (exit code 77 of the test program)."""
    from reth_b200 import _lib
    if _lib.load().b200_device_count() > 0:
        pytest.skip("a GPU is present")
    _ensure_built()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode == 77 and "no CUDA device" in r.stdout
