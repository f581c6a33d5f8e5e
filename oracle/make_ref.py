"""Populates oracle/_ref/ with the UNMODIFIED reference (linkedin/DuaLip v5.0.1) so that it travels to the GPU box.

TEST INFRASTRUCTURE / baseline only: nothing under dualip_b200/ imports from here.  oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so `bench.py --impl reference` and
`benchmark/reference_gpu_baseline.py` can time the reference's own PyTorch path on the B200 box's host cores and on its
GPU.  The reference is pure Python: "building" it is an offline `pip install --no-deps --target oracle/_ref` from a
scratch copy of /root/reference (pip writes build files into the source tree, which is read-only), falling back to a
plain copy of src/dualip.  The reference imports `mlflow` unconditionally (utils/mlflow_utils.py:5) and mlflow is not
in this image: a no-op stub package is written next to it (every reference logging call is disabled by default).

    python oracle/make_ref.py            # no-op when /root/reference is absent (the GPU box uses the prebuilt copy)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DUALIP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MLFLOW_STUB = '''"""No-op stand-in for mlflow (not installed in this image; the reference imports it unconditionally)."""


def __getattr__(name):
    return lambda *a, **k: None
'''


def available() -> bool:
    return os.path.isdir(os.path.join(OUT, "dualip"))


def make(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref/dualip exists afterwards."""
    if not os.path.isdir(os.path.join(REF_SRC, "src", "dualip")):
        return available()
    stamp = os.path.join(OUT, ".stamp")
    want = _tree_stamp(os.path.join(REF_SRC, "src", "dualip"))
    if available() and os.path.exists(stamp) and open(stamp).read() == want:
        return True
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT, exist_ok=True)
    how = "copy"
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, work, ignore=shutil.ignore_patterns(".git", "docs", "examples"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", OUT, work]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode == 0 and os.path.isdir(os.path.join(OUT, "dualip")):
            how = "pip install --no-deps --target"
        else:
            if verbose:
                print(res.stdout, res.stderr)
            shutil.rmtree(os.path.join(OUT, "dualip"), ignore_errors=True)
            shutil.copytree(os.path.join(REF_SRC, "src", "dualip"), os.path.join(OUT, "dualip"))
    # the benchmark's input generator (benchmark/generate_synthetic_data.py) is the spec of the synthetic workload
    os.makedirs(os.path.join(OUT, "reference_benchmark"), exist_ok=True)
    for name in ("generate_synthetic_data.py", "config.py", "benchmark_utils.py"):
        src = os.path.join(REF_SRC, "benchmark", name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(OUT, "reference_benchmark", name))
    os.makedirs(os.path.join(OUT, "mlflow"), exist_ok=True)
    with open(os.path.join(OUT, "mlflow", "__init__.py"), "w") as fh:
        fh.write(MLFLOW_STUB)
    with open(stamp, "w") as fh:
        fh.write(want)
    with open(os.path.join(OUT, "HOW"), "w") as fh:
        fh.write(how + "\n")
    if verbose:
        print(f"oracle/_ref populated ({how})")
    return available()


def _tree_stamp(root: str) -> str:
    items = []
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                items.append(f"{os.path.relpath(p, root)}:{os.path.getsize(p)}")
    return "\n".join(items)


def import_reference():
    """Puts oracle/_ref first on sys.path and returns the reference's `dualip` package (NOT dualip_b200's alias)."""
    if not available():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference exists")
    for name in [k for k in sys.modules if k == "dualip" or k.startswith("dualip.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "") or OUT not in (mod.__file__ or ""):
            del sys.modules[name]
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import dualip  # noqa: F401

    assert OUT in dualip.__file__, dualip.__file__
    return dualip


if __name__ == "__main__":
    ok = make(verbose=True)
    print("available" if ok else "reference tree not found and no prebuilt copy")
