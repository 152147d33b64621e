"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports exactly the
symbols include/oddio_b200.h declares; the Python binding table agrees with the header; and
without a CUDA device every entry point fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oddio_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(odb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from oddio_b200 import build, _lib

    build.build()
    return _lib.load()


def test_header_declares_entry_points():
    syms = header_symbols()
    assert "odb_scene_sample" in syms and "odb_mixer_sample" in syms and "odb_scene_run" in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header(lib):
    from oddio_b200 import _lib

    bound = set(_lib.SIGNATURES) | set(_lib.NON_STATUS)
    assert bound == set(header_symbols())


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is for CPU-only hosts")
    h = C.c_void_p()
    rc = lib.odb_ctx_create(0, C.byref(h))
    assert rc == -2  # ODB_E_CUDA
    assert b"no CPU fallback" in lib.odb_last_error()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under oddio_b200/ may reference it."""
    pkg = os.path.join(ROOT, "oddio_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in text and "oddio_oracle" not in text, f"{f} references the oracle"


def test_abi_version(lib):
    assert lib.odb_abi_version() >= 1
