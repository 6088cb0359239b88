"""The C-ABI library loads without a GPU and exports every function include/b200trie.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b200trie.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"B200_API[^;(]*?\b(b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from reth_b200 import LIB_PATH
    lib = ctypes.CDLL(LIB_PATH)
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_fails_loudly():
    """Without a CUDA device there is no fallback: b200_create returns NULL / B200_ERR_NO_DEVICE and the Python
    Engine raises."""
    from reth_b200 import _lib
    L = _lib.load()
    if L.b200_device_count() > 0:
        return  # on the GPU box this is covered by the gpu tests
    assert not L.b200_create(0)
    assert L.b200_create_status() == _lib.ERR_NO_DEVICE
    import pytest
    from reth_b200 import B200Error, Engine
    with pytest.raises(B200Error):
        Engine(0)


def test_struct_layouts_match_header():
    from reth_b200 import ACCOUNT_DTYPE
    from reth_b200._lib import FrontierEntry, Stats, Updates
    assert ACCOUNT_DTYPE.itemsize == 72
    assert ctypes.sizeof(FrontierEntry) == 68
    assert ctypes.sizeof(Stats) == 56
    assert ctypes.sizeof(Updates) == 8 * 10


def test_product_does_not_import_oracle():
    """reth_b200/ must never reach into oracle/ (the oracle is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "reth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
