"""adanet.ensemble mirror (adanet/ensemble/__init__.py:24-53)."""

from adanet_b200.ensemble.ensembler import Ensemble
from adanet_b200.ensemble.ensembler import Ensembler
from adanet_b200.ensemble.ensembler import TrainOpSpec
from adanet_b200.ensemble.mean import MeanEnsemble
from adanet_b200.ensemble.mean import MeanEnsembler
from adanet_b200.ensemble.strategy import AllStrategy
from adanet_b200.ensemble.strategy import Candidate
from adanet_b200.ensemble.strategy import GrowStrategy
from adanet_b200.ensemble.strategy import SoloStrategy
from adanet_b200.ensemble.strategy import Strategy
from adanet_b200.ensemble.weighted import ComplexityRegularized
from adanet_b200.ensemble.weighted import ComplexityRegularizedEnsembler
from adanet_b200.ensemble.weighted import MixtureWeightType
from adanet_b200.ensemble.weighted import WeightedSubnetwork

__all__ = [
    "Ensemble", "Ensembler", "TrainOpSpec", "AllStrategy", "Candidate", "GrowStrategy", "SoloStrategy", "Strategy",
    "ComplexityRegularized", "ComplexityRegularizedEnsembler", "MeanEnsemble", "MeanEnsembler", "MixtureWeightType",
    "WeightedSubnetwork",
]
