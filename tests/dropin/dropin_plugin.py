"""pytest plugin for tests/test_reference_suite_dropin.py: makes `import dualip` resolve to dualip_b200 (dualip_b200.install_as)
before the reference's own test modules are collected."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dualip_b200  # noqa: E402

dualip_b200.install_as("dualip")
