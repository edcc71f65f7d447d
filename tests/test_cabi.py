"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol that
include/biblbm_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import pathlib
import re

import pytest

import biblbm_b200 as bib

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "biblbm_b200.h").read_text()
    return sorted(set(re.findall(r"\b(bib_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = bib.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"library does not export {n}"
    assert sorted(bib.SYMBOLS) == names, "api.SYMBOLS out of sync with the header"


def test_version_and_error_string():
    lib = bib.load_library()
    assert lib.bib_version() == 100
    assert lib.bib_domain_set_option(None, 1, C.c_double(0.0)) == 0
    assert b"null domain" in lib.bib_last_error()


def test_bad_arguments_return_zero():
    lib = bib.load_library()
    h = C.c_void_p()
    per = (C.c_int * 3)(1, 1, 1)
    assert lib.bib_domain_create(0, 4, 4, 4, C.c_double(1.0), C.c_double(1.0), per, 0, C.byref(h)) == 0
    assert lib.bib_domain_create(4, 4, 4, 5, C.c_double(1.0), C.c_double(1.0), per, 0, C.byref(h)) == 0      # n_ext is 3 or 4
    assert lib.bib_domain_create(4, 4, 4, 4, C.c_double(2.5), C.c_double(1.0), per, 0, C.byref(h)) == 0
    assert not h


def test_default_inplace_argument_checking():
    """bib_set_default_inplace: -1 (automatic), 0 (two buffers) or a shift of >= 2 planes; nothing else (a shift of 1 leaves no plane per launch)."""
    lib = bib.load_library()
    try:
        for ok in (-1, 0, 2, 17, 1000):
            assert lib.bib_set_default_inplace(ok) == 1
        for bad in (1, -2, -100):
            assert lib.bib_set_default_inplace(bad) == 0
            assert b"set_default_inplace" in lib.bib_last_error()
    finally:
        lib.bib_set_default_inplace(-1)


def test_null_handle_is_rejected_by_every_entry_point():
    """Every entry point that takes a domain handle returns 0 (never crashes) for a NULL handle and names the problem."""
    lib = bib.load_library()
    text = (ROOT / "include" / "biblbm_b200.h").read_text()
    protos = re.findall(r"^int (bib_[a-z0-9_]+)\(bib_domain\* d([^)]*)\);", text, flags=re.M)
    assert len(protos) >= 45
    for name, rest in protos:
        if name == "bib_domain_destroy":          # like free(NULL): nothing to do, succeeds
            assert lib.bib_domain_destroy(None) == 1
            continue
        n_args = 0 if not rest.strip() else rest.count(",")
        args = []
        for a in [x.strip() for x in rest.split(",")[1:]]:
            args.append(C.c_double(0.0) if a.startswith("double ") else (C.c_longlong(0) if a.startswith("long long ") else (C.c_void_p(None) if "*" in a or "[" in a else C.c_int(0))))
        assert len(args) == n_args
        assert getattr(lib, name)(None, *args) == 0, name
        assert b"null domain" in lib.bib_last_error(), name


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly instead of computing on the host."""
    if bib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bib.BibError, match="no CUDA device"):
        bib.Lattice(8, 8, 8, 1.0, 0.5)


def test_no_cpu_fallback_multi_slab():
    """The multi-slab handle (bib_domain_create_multi, what the overlay's BIB_GPUS=N uses) fails just as loudly without a GPU."""
    if bib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bib.BibError, match="no CUDA device"):
        bib.Lattice(48, 8, 8, 1.0, 0.5, slabs=2, devices=[0, 0])


def test_built_for_sm100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(bib.api.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
