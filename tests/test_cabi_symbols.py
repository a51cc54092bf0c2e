"""The C-ABI shared object loads on a GPU-less machine, exports every symbol include/mmgclip_b200.h declares, the
ctypes table mirrors the header one to one, and calls fail loudly (no CPU fallback).  No compute is launched."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from mmgclip_b200 import _lib

HEADER = os.path.join(ROOT, "include", "mmgclip_b200.h")


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|long long|const char\s*\*)\s+(mmg_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    protos = header_prototypes()
    assert len(protos) >= 25
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mmg_\w+)", out))
    assert set(protos) <= exported
    assert exported <= set(protos), f"exported but undeclared: {exported - set(protos)}"


def test_ctypes_table_matches_header():
    protos = header_prototypes()
    assert set(_lib.SIGNATURES) == set(protos)
    for name, n_args in protos.items():
        assert len(_lib.SIGNATURES[name][1]) == n_args, name


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.mmg_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_only_sm100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_calls_fail_loudly_without_gpu_or_with_host_pointers():
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MmgError):
            _lib.device_info()
    buf = (ctypes.c_float * 16)()
    rc = lib.mmg_cast_f32_to_bf16(ctypes.addressof(buf), ctypes.addressof(buf), 16, None)
    assert rc < 0 and _lib.last_error()
    with pytest.raises((ValueError, _lib.MmgError)):
        _lib.check(rc, "mmg_cast_f32_to_bf16")
    rc = lib.mmg_gemm(1, None, 8, 0, None, 8, 0, None, 8, 0, 8, 8, 1.0, None, None, 0, 0, 1, None)
    assert rc == -1  # empty problem -> MMG_ERR_BAD_ARG


def test_missing_library_is_an_import_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmmgclip_b200.so")
    with pytest.raises(ImportError, match="no CPU"):
        _lib.load()
