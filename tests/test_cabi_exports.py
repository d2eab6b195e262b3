"""The C-ABI library loads on a CPU-only box and exports every symbol include/caphn_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "caphn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\bint\s+(caphn_\w+)\s*\(", src)


def test_library_builds_and_exports_every_declared_symbol():
    from hypernet_image_captioning_b200 import _cabi, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/caphn_b200.h but not exported"
    # the ctypes table and the header must describe the same set of functions
    assert set(names) == set(_cabi.SIGNATURES), set(names) ^ set(_cabi.SIGNATURES)


def test_header_arity_matches_ctypes_table():
    from hypernet_image_captioning_b200 import _cabi
    src = open(os.path.join(ROOT, "include", "caphn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\bint\s+(caphn_\w+)\s*\(([^)]*)\)", src):
        assert len(args.split(",")) == len(_cabi.SIGNATURES[name]), name


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from hypernet_image_captioning_b200 import ops, _cabi
    with pytest.raises(_cabi.CaphnError):
        ops.linear(torch.zeros(2, 2), torch.zeros(2, 2))
