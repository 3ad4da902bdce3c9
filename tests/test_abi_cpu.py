"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/hevcdl.h
declares, and refuses to run without a GPU (no CPU fallback on the product path)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared(header):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(hevcdl_[a-z0-9_]+)\s*\(", hdr))


def test_library_exports_every_declared_symbol(built, host):
    public, internal = _declared("hevcdl.h"), _declared("hevcdl_internal.h")
    assert public == set(host.EXPORTS), public ^ set(host.EXPORTS)
    assert internal == set(host.EXPORTS_INTERNAL), internal ^ set(host.EXPORTS_INTERNAL)
    assert not any("bench" in n or "debug" in n for n in public)      # measurement / test hooks stay out of the boundary
    lib = C.CDLL(built)
    for name in public | internal:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_header(host):
    assert C.sizeof(host.Cfg) == 12 * 4 + 8                 # 12 int32, the weights pointer
    hdr = open(os.path.join(ROOT, "include", "hevcdl.h")).read()
    assert int(re.search(r"#define HEVCDL_ABI_VERSION (\d+)", hdr).group(1)) == host.ABI_VERSION
    assert host.PU_DTYPE.itemsize == 8
    assert C.sizeof(host.Stats) == 48


def test_create_fails_loudly_without_gpu(built, host):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(host.HevcdlError) as e:
        host.DepthPredictor(416, 240)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_cfg(built, host):
    lib = host.load_library()
    h = C.c_void_p()
    cfg = host.Cfg(host.ABI_VERSION, 0, 417, 240, 1, 0, 1, 0, 1, 0, 0, 0, b"x")       # width not a multiple of 8
    assert lib.hevcdl_create(C.byref(cfg), C.byref(h)) == -1
    cfg = host.Cfg(99, 0, 416, 240, 1, 0, 1, 0, 1, 0, 0, 0, b"x")      # wrong ABI version
    assert lib.hevcdl_create(C.byref(cfg), C.byref(h)) == -1
    cfg = host.Cfg(host.ABI_VERSION, 0, 416, 240, 1, 0, 1, 0, 1, 8, 0, 0, b"x")       # unknown output flag
    assert lib.hevcdl_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.hevcdl_status_str(-2).decode().startswith("no usable CUDA device")


def test_product_code_never_imports_oracle():
    pk = os.path.join(ROOT, "hevc-deep-learning-pipeline_b200")
    for d, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(d, f)
    for f in os.listdir(os.path.join(ROOT, "hm_plugin")) if os.path.isdir(os.path.join(ROOT, "hm_plugin")) else []:
        if f.endswith((".cpp", ".h")):
            assert "liboracle" not in open(os.path.join(ROOT, "hm_plugin", f)).read()
