"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/gempic_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gempic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(gempic_[a-z0-9_]+)\s*\(", text))
    names.discard("gempic_func1d")
    names.discard("gempic_func2d")
    return sorted(names)


def test_header_declares_the_boundary():
    names = declared_symbols()
    assert len(names) >= 60
    for must in ("gempic_hs_strang_splitting", "gempic_pg_upload", "gempic_pmc1d_add_charge",
                 "gempic_maxwell1d_compute_e_from_j", "gempic_boris_strang_splitting", "gempic_diag_write_step",
                 "gempic_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol(gp):
    lib = gp.load()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"libgempic_b200.so lacks {missing}"


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "gempic_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments may mention torch.distributed
    assert "torch" not in code.lower() and "at::" not in code and "#include <torch" not in text


def test_fails_loudly_without_gpu(gp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = gp.load()
    rc = lib.gempic_init(ctypes.c_int(0))
    assert rc == 6  # GEMPIC_ENOTINIT
    assert b"no CUDA device" in lib.gempic_last_error() or b"CPU path" in lib.gempic_last_error()
    h = ctypes.c_uint64(0)
    rc = lib.gempic_pg_create(ctypes.c_int(1), ctypes.c_int(2), ctypes.c_int(1), ctypes.c_int64(4), ctypes.c_double(1.0),
                              ctypes.c_double(1.0), ctypes.c_double(0.0), ctypes.byref(h))
    assert rc == 6
    with pytest.raises(gp.GempicError):
        gp.ParticleGroup(1, 2, 4)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gempic.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} references the oracle"
