"""Drop-in check of the host-side API: the reference's OWN test modules, unmodified, collected from /root/reference/tests and run
against this package through the `dualip` alias (dualip_b200.install_as).  Only the modules that need no computation on CPU
tensors can pass here -- objectives, projections and preprocessing are CUDA-only by design and their reference tests use
CPU tensors.  Skipped where the reference tree is absent (the GPU box)."""
import os
import subprocess
import sys

import pytest

REF_TESTS = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    ("test_utils.py", "", 5),      # agd_utils: norm_of_difference, history ring, Lipschitz estimate, step-size rule
    ("test_agd.py", "", 2),        # AcceleratedGradientDescent on Python objectives, incl. the four known-answer trace values
    ("test_import.py", "", 1),
    ("test_equality_constraints.py", "test_project_on_nn_cone", 1),
    ("preprocessing/test_input_validation.py", "", None),
    # setup-time CSC helpers (index arithmetic, any device); the per-iteration operators of that module are CUDA kernels here and
    # are covered by tests/test_gpu_operators.py with the same cases on device tensors
    ("test_sparse_utils.py", "vstack or hstack or combined or right_multiply", 4),  # vectorised checks, same errors as the reference's per-column loop
]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not present")
@pytest.mark.parametrize("module,select,expected", CASES)
def test_reference_test_module_passes_against_this_package(module, select, expected):
    env = dict(os.environ, PYTHONPATH=os.path.join(HERE, "dropin") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "dropin_plugin", "-p", "no:cacheprovider", "-c", os.devnull,
           os.path.join(REF_TESTS, module)]
    if select:
        cmd += ["-k", select]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=os.path.join(HERE, "dropin"), timeout=600)
    tail = res.stdout[-1500:] + res.stderr[-1500:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout and "failed" not in res.stdout, tail
    if expected is not None:
        assert f"{expected} passed" in res.stdout, tail
