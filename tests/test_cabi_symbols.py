"""The C-ABI library loads and exports every symbol include/gdrn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gdrn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25, names
    for n in names:
        assert hasattr(lib, n), f"libgdrn_b200.so does not export {n}"


def test_python_binding_covers_header(lib):
    from gdrnpp_bop2022_b200 import _lib

    assert set(declared_symbols()) == set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_channel(lib):
    assert lib.gdrn_version() >= 100
    assert isinstance(lib.gdrn_last_error(), bytes)
    assert lib.gdrn_launch_count() >= 0


def test_no_torch_types_in_abi():
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gdrn_b200.h")).read(), flags=re.S)
    assert "at::" not in src and "torch" not in src and "Tensor" not in src
