"""CPU-only: the C-ABI library loads and exports every symbol include/mcmcdiag_b200.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import mcmcdiag_b200 as m
    return m._lib.load()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mcmcdiag_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    import mcmcdiag_b200 as m
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in m._lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(m._lib.SIGNATURES) == syms


def test_abi_version(lib):
    assert lib.mcd_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.mcd_create(ctypes.byref(h), 0)
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.mcd_create_error()
    import mcmcdiag_b200 as m
    import numpy as np
    with pytest.raises(m._lib.MCDLibraryError):
        m.ess_rhat(np.zeros((100, 4, 2)))


def test_host_side_argument_errors_need_no_gpu():
    import numpy as np
    import mcmcdiag_b200 as m
    x = np.zeros((100, 4, 2))
    with pytest.raises(m.ArgumentError):
        m.ess_rhat(x, kind="foo")
    with pytest.raises(m.ArgumentError):
        m.rhat(x, kind="foo")
    with pytest.raises(m.DimensionMismatch):
        m.rhat_nested(x, [1, 2, 3])
    with pytest.raises(m.ArgumentError):
        m.rhat_nested(x, [1, 1, 1, 1])
    with pytest.raises(m.ArgumentError):
        m.rhat_nested(x, [1, 1, 1, 2])
    inds = m.api._validate_superchain_ids(["b", "a", "b", "a"], 4)
    assert inds.tolist() == [[1, 0], [3, 2]]


def _header_prototypes():
    """name -> list of parameter C types parsed from include/mcmcdiag_b200.h"""
    hdr = open(os.path.join(ROOT, "include", "mcmcdiag_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(mcd_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        if args == ["void"] or args == [""]:
            args = []
        protos[m.group(2)] = args
    return protos


def _kind(ctype_decl):
    """Coarse class of a C parameter declaration."""
    d = ctype_decl
    if "*" in d:
        return "ptr"
    if "double" in d:
        return "double"
    if "int64_t" in d or "uint64_t" in d or "long long" in d:
        return "i64"
    return "int"       # int, unsigned


def test_ctypes_signatures_match_header_prototypes():
    """Every ctypes signature has the header's arity and the same coarse type per argument, so that a
    drifted binding is caught without a GPU."""
    import ctypes as C
    import mcmcdiag_b200 as m
    protos = _header_prototypes()
    assert set(protos) == set(m._lib.SIGNATURES)

    def ckind(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        if t is C.c_double:
            return "double"
        if t in (C.c_int64, C.c_uint64):
            return "i64"
        return "int"

    for name, (_, argtypes) in m._lib.SIGNATURES.items():
        want = [_kind(a) for a in protos[name]]
        got = [ckind(t) for t in argtypes]
        assert got == want, (name, got, want, protos[name])
