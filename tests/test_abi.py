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


def test_build_is_up_to_date_by_source_digest(monkeypatch, tmp_path):
    """build() compares a digest of the sources with the one recorded beside the library, not file times: a copied tree
    (the snapshot on a GPU box) must not recompile, an edited source must."""
    import subprocess

    stamp = os.path.join(os.path.dirname(_native.LIB_PATH), "SOURCE_DIGEST")
    if not os.path.exists(stamp):
        pytest.skip("library was built by make directly (no digest recorded)")
    assert open(stamp).read().strip() == _native._source_digest(), "libdualip_b200.so is stale: run __graft_entry__.build()"

    def no_make(*a, **k):
        raise AssertionError("make must not run when the digest matches")

    monkeypatch.setattr(subprocess, "run", no_make)
    assert _native.build() == _native.LIB_PATH
    # a changed source changes the digest
    csrc = tmp_path / "pkg" / "csrc"  # the header sits at <package>/../include, as in the repository
    csrc.mkdir(parents=True)
    (tmp_path / "include").mkdir()
    for name in os.listdir(_native.CSRC_DIR):
        if name.endswith((".cu", ".cuh", ".inc")) or name == "Makefile":
            (csrc / name).write_bytes(open(os.path.join(_native.CSRC_DIR, name), "rb").read())
    (tmp_path / "include" / "dualip_b200.h").write_bytes(open(os.path.join(ROOT, "include", "dualip_b200.h"), "rb").read())
    monkeypatch.setattr(_native, "CSRC_DIR", str(csrc))
    same = _native._source_digest()
    (csrc / "lp.cu").write_bytes((csrc / "lp.cu").read_bytes() + b"\n// edited\n")
    assert _native._source_digest() != same
