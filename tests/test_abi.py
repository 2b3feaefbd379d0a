"""The C-ABI library loads and exports every entry point that include/fiat_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fiat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fiatb200_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("fiatb200_simplex_plan_create", "fiatb200_tensor_plan_create", "fiatb200_lattice_plan_create",
                 "fiatb200_tabulate", "fiatb200_tabulate_host", "fiatb200_locate_subcells", "fiatb200_plan_destroy"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from fiat_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert set(_lib.EXPORTS) == set(declared_functions())
    assert lib.fiatb200_version() == 1


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the header structs: sizes follow from the field lists in the header."""
    from fiat_b200 import _lib
    assert ctypes.sizeof(_lib.EntityMapStruct) == 4 + 4 + 9 * 8 + 3 * 8
    assert ctypes.sizeof(_lib.TensorLeafStruct) == 8 + ctypes.sizeof(_lib.EntityMapStruct) + 8
    # the program struct is all int32 / pointer / int64 fields; its size is a multiple of 8
    assert ctypes.sizeof(_lib.SimplexProgramStruct) % 8 == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from fiat_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except _lib.LibraryError as exc:
        assert "no CPU fallback" in str(exc)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
