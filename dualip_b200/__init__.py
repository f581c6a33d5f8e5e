"""dualip_b200: B200-native implementation of DuaLip's per-iteration dual-ascent hot path.

Module layout mirrors the reference package (`dualip.run_solver`, `dualip.types`, `dualip.objectives.matching`,
`dualip.projections`, `dualip.optimizers.agd`, `dualip.utils.dist_utils`, `dualip.preprocessing.precondition`), so
`import dualip_b200 as dualip` — or `dualip_b200.install_as("dualip")` for code that imports submodules — is a drop-in
for that path.  The compute lives in dualip_b200/csrc (CUDA, sm_100a) behind include/dualip_b200.h."""
import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = [
    "types", "run_solver", "objectives", "objectives.base", "objectives.matching", "objectives.miplib", "objectives.matching_fairness",
    "projections", "projections.base",
    "projections.box", "projections.cone", "projections.simplex", "optimizers", "optimizers.agd", "optimizers.agd_utils",
    "utils", "utils.dist_utils", "utils.sparse_utils", "utils.mlflow_utils", "utils.step_size_utility", "preprocessing", "preprocessing.precondition",
    "preprocessing.input_validation",
]


def install_as(alias: str = "dualip") -> None:
    """Register this package and its submodules under another top-level name in sys.modules."""
    pkg = sys.modules[__name__]
    sys.modules[alias] = pkg
    for sub in _SUBMODULES:
        sys.modules[f"{alias}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
