"""No-GPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/arvc_icp.h declares, struct layouts match, and the engine fails loudly without a device."""
import ctypes
import os
import re

import pytest

from lidar_slam_arvc_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _built()
    header = open(os.path.join(ROOT, "include", "arvc_icp.h")).read()
    declared = set(re.findall(r"\b(arvc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(engine.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(engine.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.arvc_version() >= 100


def test_struct_layouts():
    assert ctypes.sizeof(engine.ResultRecord) == 160
    assert ctypes.sizeof(engine.IcpParams) == 32
    assert ctypes.sizeof(engine.PreprocessParams) == 72
    assert engine.RESULT_DTYPE.fields["T"][1] == 16 and engine.RESULT_DTYPE.fields["fitness"][1] == 144


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _built()
    with pytest.raises(engine.EngineError, match="no CUDA device|CUDA"):
        engine.Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lidar_slam_arvc_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lidar_slam_arvc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src and "icp_oracle" not in src, os.path.join(dirpath, f)
