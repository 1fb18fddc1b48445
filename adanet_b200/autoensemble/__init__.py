"""adanet.autoensemble mirror (adanet/autoensemble/__init__.py)."""

from adanet_b200.autoensemble.common import AutoEnsembleSubestimator
from adanet_b200.autoensemble.estimator import AutoEnsembleEstimator

__all__ = ["AutoEnsembleEstimator", "AutoEnsembleSubestimator"]
