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


def test_init_devices_fails_loudly_without_gpu(gp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = gp.load()
    ids = (ctypes.c_int * 2)(0, 1)
    assert lib.gempic_init_devices(ctypes.c_int(2), ids) == 6        # GEMPIC_ENOTINIT: no CPU path in this mode either
    assert lib.gempic_device_count() == 1
    with pytest.raises(gp.GempicError):
        gp.init_devices(2)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gempic.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} references the oracle"


def _c_prototypes():
    """name -> number of parameters, from include/gempic_b200.h"""
    text = open(os.path.join(ROOT, "include", "gempic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(gempic_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        if name in ("gempic_func1d", "gempic_func2d"):
            continue
        protos[name] = 0 if params in ("", "void") else len(_split_top(params))
    return protos


def _split_top(s):
    """split on commas that are not inside parentheses / brackets"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def _julia_ccalls():
    """(symbol, number of argument types, number of arguments) of every ccall in julia/GEMPICB200.jl"""
    text = open(os.path.join(ROOT, "julia", "GEMPICB200.jl")).read()
    text = re.sub(r"#[^\n]*", "", text)
    out = []
    for m in re.finditer(r"ccall\(\(:(gempic_[a-z0-9_]+),\s*LIB\)\s*,", text):
        i, depth = m.end(), 1            # scan to the parenthesis that closes ccall(
        j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        args = _split_top(text[i:j - 1])    # return type, (argument types...), arguments...
        types = args[1].strip()
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        inner = types[1:-1].strip().rstrip(",")
        n_types = 0 if not inner else len(_split_top(inner))
        out.append((m.group(1), n_types, len(args) - 2))
    return out


def test_julia_shim_binds_the_declared_prototypes():
    """Julia is not installed here, so the shim is checked statically: every ccall names an entry point declared in the
    header, with as many argument types -- and as many arguments -- as the C prototype has parameters."""
    protos = _c_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 40
    for name, n_types, n_args in calls:
        assert name in protos, f"{name} is not declared in include/gempic_b200.h"
        assert n_types == protos[name], f"{name}: {n_types} ccall argument types, {protos[name]} C parameters"
        assert n_args == n_types, f"{name}: {n_args} arguments for {n_types} argument types"


def _c_param_class(p):
    p = re.sub(r"\bconst\b", "", p).strip()
    stars = p.count("*")
    base = re.sub(r"[\*\s]+[A-Za-z_0-9\[\]]*$", "", p.replace("*", " * ")).strip() if stars else p.rsplit(None, 1)[0].strip()
    base = re.sub(r"\s+", " ", base.replace("*", "")).strip()
    if base in ("gempic_func1d", "gempic_func2d"):
        return "ptr"
    if base == "gempic_handle":
        return "handle*" if stars else "handle"
    if base == "double":
        return "double*" if stars else "double"
    if base == "int":
        return "int*" if stars else "int"
    if base == "int64_t":
        return "int64*" if stars else "int64"
    if base == "uint64_t":
        return "uint64*" if stars else "uint64"
    if base in ("void", "char"):
        return "ptr"
    return base + ("*" * stars)


_JL = {"Handle": "handle", "UInt64": "handle", "Cint": "int", "Int64": "int64", "Cdouble": "double", "Cstring": "ptr",
       "Ref{Handle}": "handle*", "Ptr{Handle}": "handle*", "Ptr{Cdouble}": "double*", "Ref{Cdouble}": "double*",
       "Ptr{Cvoid}": "ptr", "Ptr{UInt8}": "ptr", "Ref{Cint}": "int*", "Ptr{Cint}": "int*", "Ref{Int64}": "int64*",
       "Ptr{Int64}": "int64*", "Ptr{Ptr{Cdouble}}": "double*", "Ref{Ptr{Cdouble}}": "double*"}


def test_julia_shim_argument_types_match_the_header():
    """same static check, type by type (handle, int, int64, double and their pointers)"""
    htext = open(os.path.join(ROOT, "include", "gempic_b200.h")).read()
    htext = re.sub(r"/\*.*?\*/", "", htext, flags=re.S)
    cparams = {}
    for m in re.finditer(r"\b(gempic_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", htext, flags=re.S):
        ps = m.group(2).strip()
        cparams[m.group(1)] = [] if ps in ("", "void") else [_c_param_class(p) for p in _split_top(ps)]
    jtext = re.sub(r"#[^\n]*", "", open(os.path.join(ROOT, "julia", "GEMPICB200.jl")).read())
    checked = 0
    for m in re.finditer(r"ccall\(\(:(gempic_[a-z0-9_]+),\s*LIB\)\s*,", jtext):
        i, depth = m.end(), 1
        j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(jtext[j], 0)
            j += 1
        types = _split_top(jtext[i:j - 1])[1].strip()[1:-1].strip().rstrip(",")
        jl = [t.strip() for t in _split_top(types)] if types else []
        want = cparams[m.group(1)]
        for k, (jt, ct) in enumerate(zip(jl, want)):
            assert jt in _JL, f"{m.group(1)} argument {k}: unknown Julia type {jt}"
            got = _JL[jt]
            ok = (got == ct or (ct == "uint64" and got == "handle") or (ct == "uint64*" and got == "handle*")
                  or (ct == "double**" and got == "double*"))
            assert ok, f"{m.group(1)} argument {k}: Julia {jt} vs C {ct}"
            checked += 1
    assert checked > 150


def test_julia_shim_keeps_pg_array_coherent():
    """`pg.array` is a lazily synchronised host mirror in the shim: every shim function that lets the device change the
    particles flushes a written mirror first and marks it stale afterwards (static check: Julia is not installed here)"""
    text = re.sub(r"#[^\n]*", "", open(os.path.join(ROOT, "julia", "GEMPICB200.jl")).read())
    assert "function Base.getproperty(pg::ParticleGroup, name::Symbol)" in text
    assert "function Base.setproperty!(pg::ParticleGroup, name::Symbol, value)" in text
    movers = ("gempic_hs_operator_host", "gempic_hs_strang_splitting_host", "gempic_boris_staggering_host",
              "gempic_boris_strang_splitting_host", "gempic_hs2d_operator_host", "gempic_hs2d_strang_splitting_host", "gempic_pg_sort")
    readers = ("gempic_solve_poisson", "gempic_diag_write_step", "gempic_hs2d_charge_density", "gempic_pmc1d_add_charge_pg")
    # split into top-level function bodies ("function ... end" blocks at column 0)
    blocks = re.findall(r"^function .*?^end", text, flags=re.S | re.M)
    for sym in movers + readers:
        body = [b for b in blocks if f":{sym}," in b]
        assert body, f"{sym} is not called from a function block"
        for b in body:
            assert "_flush(" in b, f"{sym}: the written host mirror is not uploaded first"
            if sym in movers:
                assert "_touched(" in b, f"{sym}: the host mirror is not marked stale afterwards"


def test_julia_shim_blocks_balance():
    """coarse syntax check without Julia: block openers (function, struct, module, for, while, try, let, begin, if, do) and
    `end`s balance once comments, strings and bracketed index expressions (`a[1:end]`) are stripped"""
    text = open(os.path.join(ROOT, "julia", "GEMPICB200.jl")).read()
    text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
    text = re.sub(r"#[^\n]*", "", text)
    prev = None
    while prev != text:
        prev = text
        text = re.sub(r"\[[^\[\]]*\]", "0", text)     # innermost first; the placeholder has no brackets, so nesting resolves
    openers = closers = 0
    for ln in text.splitlines():
        for kw in ("function", "module", "for", "while", "try", "let", "begin", "quote", "macro"):
            openers += len(re.findall(r"(?<![A-Za-z0-9_!.:])" + kw + r"(?![A-Za-z0-9_!])", ln))
        openers += len(re.findall(r"(?<![A-Za-z0-9_!.])struct(?![A-Za-z0-9_!])", ln))
        openers += len(re.findall(r"\babstract type\b", ln))
        openers += len(re.findall(r"(?<![A-Za-z0-9_!])if\s", ln)) - len(re.findall(r"elseif\s", ln))
        openers += 1 if re.search(r"\bdo(\s+[a-z_, ]+)?\s*$", ln.strip()) else 0
        closers += len(re.findall(r"(?<![A-Za-z0-9_!:])end(?![A-Za-z0-9_!])", ln))
    assert openers == closers and openers > 30, (openers, closers)


# header symbols the Julia shim deliberately does not bind, with the reason
JULIA_UNBOUND = {
    "gempic_version": "informational",
    "gempic_stream": "raw cudaStream_t for CUDA-event timing (bench.py)",
    "gempic_device_info": "bench / tests",
    "gempic_launch_count": "bench / tests",
    "gempic_profile_enable": "bench (per-pass device timing)",
    "gempic_profile_read": "bench (per-pass device timing)",
    "gempic_set_option": "no option is defined in this version",
    "gempic_comm_allreduce": "test hook of the all-reduce path",
    "gempic_sobol_points": "test hook of the Sobol direction numbers",
    "gempic_pg_row_ptr": "raw device pointers (torch tensors in the Python tests); a Julia caller uses pg.array / upload!",
    "gempic_pg_set_row_device": "raw device pointers",
    "gempic_pg_get_row_device": "raw device pointers",
    "gempic_pg_info": "the shim struct keeps dims, n_particles, charge, mass, common_weight itself",
    "gempic_maxwell1d_l2norm_squared": "l2norm_squared(m, c, degree) = inner_product(m, c, c, degree) in the shim, as in the reference",
    "gempic_maxwell1d_get_table": "constructor tables, read by the parity tests only",
    "gempic_maxwell2d_get_table": "constructor tables, read by the parity tests only",
    "gempic_maxwell2d_solve_mass": "internal solver step exposed for the parity tests",
    "gempic_maxwell2d_multiply_mass": "internal solver step exposed for the parity tests",
}


def test_every_header_symbol_is_bound_by_the_julia_shim_or_listed():
    declared = set(declared_symbols())
    bound = {name for name, _, _ in _julia_ccalls()}
    unbound = declared - bound
    assert unbound == set(JULIA_UNBOUND), (sorted(unbound - set(JULIA_UNBOUND)), sorted(set(JULIA_UNBOUND) - unbound))
