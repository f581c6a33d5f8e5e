"""The C-ABI library loads and exports every symbol include/dualip_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest

from conftest import ROOT
from dualip_b200 import _native


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dualip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dualip_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() / make -C dualip_b200/csrc"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} missing from libdualip_b200.so"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in dualip_b200/_native.py"
    assert set(_native.SIGNATURES) == set(declared)


def test_abi_version_and_error_string():
    lib = _native.lib()
    assert lib.dualip_abi_version() == 2
    assert isinstance(lib.dualip_last_error(), bytes)


def test_error_codes_map_to_reference_exception_types():
    lib = _native.lib()
    import ctypes

    rc = lib.dualip_plan_create(ctypes.byref(ctypes.c_void_p()), None)  # null descriptor: argument check only
    assert rc == _native.EINVAL
    with pytest.raises(ValueError):
        _native.check(rc, "dualip_plan_create")
    assert ctypes.sizeof(_native.ProjClass) == 24 and ctypes.sizeof(_native.Scalars) == 64


def test_sass_is_sm100a_and_uses_bulk_async_copy():
    """The library carries sm_100a code, and the hot kernel stages lambda with the TMA engine (UBLKCP) and flushes the
    gradient with a bulk reduction (UBLKRED)."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "UBLKRED" in sass and "SYNCS" in sass
