"""The C-ABI library loads and exports every symbol include/ddpmir.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ddpmir.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddpmir_[a-z0-9_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as ge
    ge.build()
    from ddpm_image_restoration_b200 import _lib
    names = header_functions()
    assert len(names) >= 25
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in ddpmir.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table out of sync with the header"
    assert _lib.lib().ddpmir_version() >= 100


def test_epilogue_struct_matches_header():
    from ddpm_image_restoration_b200 import _lib
    src = open(os.path.join(ROOT, "include", "ddpmir.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} ddpmir_epilogue_t;", src, flags=re.S).group(1)
    fields = re.findall(r"(\w+);", body)
    assert fields == [f[0] for f in _lib.Epilogue._fields_]


def test_missing_library_fails_loudly(monkeypatch):
    from ddpm_image_restoration_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libddpmir.so")
    with pytest.raises(_lib.DdpmirError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200 import ops
    from ddpm_image_restoration_b200._lib import DdpmirError
    with pytest.raises(DdpmirError):
        ops.maxpool2(torch.zeros(1, 2, 2, 8))
    with pytest.raises(RuntimeError):
        P.WebPDiffusionModel().eval()(torch.zeros(1, 3, 32, 32), torch.zeros(1))
    with pytest.raises(RuntimeError):
        P.DDRMWebPSampler(P.WebPDiffusionModel()).sample(torch.zeros(1, 3, 32, 32), 10, steps=2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ddpm_image_restoration_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
