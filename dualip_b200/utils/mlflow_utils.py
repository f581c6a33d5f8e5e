"""MLflow is observability, not part of the dual-ascent path, and is out of scope.  Only the
configuration type survives so that `run_solver(..., mlflow_config=...)` keeps the reference's signature
(src/dualip/utils/mlflow_utils.py:11-22); enabling it is an error rather than a silent no-op."""
from dataclasses import dataclass


@dataclass
class MLflowConfig:
    enabled: bool
    tracking_uri: str = ""
    experiment_name: str = ""
    run_name: str = ""
    log_hyperparameters: bool = True
    log_metrics: bool = True
    synchronous: bool = False
