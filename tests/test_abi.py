"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares (no compute calls here)."""
import ctypes
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def test_library_exports_all_header_symbols():
    import amrex_b200
    from __graft_entry__ import header_symbols
    lib = amrex_b200.load_library()
    names = header_symbols()
    assert len(names) > 100
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the reference's own C ABI names for this path are all there (Src/F_Interfaces/LinearSolvers, Base)
    for n in ("amrex_fi_new_multigrid", "amrex_fi_multigrid_solve", "amrex_fi_new_abeclaplacian", "amrex_fi_new_poisson",
              "amrex_fi_abeclap_set_scalars", "amrex_fi_abeclap_set_acoeffs", "amrex_fi_abeclap_set_bcoeffs",
              "amrex_fi_linop_set_domain_bc", "amrex_fi_linop_set_level_bc", "amrex_fi_linop_set_maxorder",
              "amrex_fi_multifab_fill_boundary", "amrex_fi_multifab_parallelcopy", "amrex_fi_new_multifab",
              "amrex_fi_new_boxarray", "amrex_fi_boxarray_maxsize", "amrex_fi_new_distromap", "amrex_fi_new_geometry"):
        assert n in names


def test_version_string():
    import amrex_b200
    lib = amrex_b200.load_library()
    lib.b200mg_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.b200mg_version()


def test_no_cpu_fallback_without_gpu():
    """Device operations must fail loudly (error through the C ABI), never fall back to a host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import amrex_b200 as ab
    with pytest.raises(ab.AmrexError):
        ab.init(0)


def test_product_does_not_reference_oracle():
    """Nothing under amrex_b200/ may import, link or execute anything under oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "amrex_b200")):
        if "lib" in root.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".H", ".h", "Makefile")):
                s = open(os.path.join(root, f), errors="ignore").read()
                if "oracle/" in s or "ref_driver" in s or "/root/reference" in s:
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def _header_prototypes():
    """name -> number of parameters, for every function declared in include/*.h."""
    import re
    protos = {}
    inc = os.path.join(REPO, "include")
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        src = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, h)).read(), flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b((?:amrex_fi|amrex_b200|b200mg)_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
            args = m.group(2).strip()
            protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_ctypes_signatures_match_the_headers():
    """Every ctypes signature of amrex_b200/capi.py has the parameter count of its C prototype: a changed C entry point
    cannot silently keep a stale Python binding."""
    from amrex_b200 import capi
    protos = _header_prototypes()
    assert len(protos) > 150
    stale = {n: (len(sig[1]), protos[n]) for n, sig in capi._SIGS.items() if n in protos and len(sig[1]) != protos[n]}
    unknown = [n for n in capi._SIGS if n not in protos]
    assert not stale, stale
    assert not unknown, unknown
