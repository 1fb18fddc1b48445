"""ComplexityRegularizedEnsembler: the AdaNet objective as an Ensembler.

API mirror of adanet/ensemble/weighted.py (WeightedSubnetwork :42-87,
ComplexityRegularized :90-136, MixtureWeightType :139-147,
ComplexityRegularizedEnsembler :150-617).  The arithmetic --

  F(w) = (1/m) sum_i Phi(sum_j w_j h_j(x_i), y_i) + sum_j (lambda r(h_j) + beta) |w_j|

-- is executed per step by the fused CUDA head kernel `adn_ensemble_head`
(csrc/heads.cu); this class holds the configuration, builds the (symbolic)
ensemble description the Estimator lowers, and exposes the small host-side
helpers the reference's unit tests pin (default initial weights, gamma,
complexity regularisation, norms/fractions).
"""

from __future__ import annotations

import collections
from typing import List, Optional, Sequence

import numpy as np

from adanet_b200 import train
from adanet_b200.ensemble.ensembler import Ensemble
from adanet_b200.ensemble.ensembler import Ensembler


class WeightedSubnetwork(collections.namedtuple(
    "WeightedSubnetwork", ["name", "iteration_number", "weight", "logits", "subnetwork"])):
  """A weight applied to a subnetwork's logits / last layer (weighted.py:42-87)."""

  def __new__(cls, name="", iteration_number=0, weight=None, logits=None, subnetwork=None):
    return super(WeightedSubnetwork, cls).__new__(cls, name=name, iteration_number=iteration_number, weight=weight,
                                                  logits=logits, subnetwork=subnetwork)


class ComplexityRegularized(collections.namedtuple(
    "ComplexityRegularized", ["weighted_subnetworks", "bias", "logits", "subnetworks", "complexity_regularization"]),
                            Ensemble):
  """F(x) = sum_i w_i h_i(x) + b, regularised by model complexity (weighted.py:90-136)."""

  def __new__(cls, weighted_subnetworks, bias, logits, subnetworks=None, complexity_regularization=None):
    return super(ComplexityRegularized, cls).__new__(
        cls, weighted_subnetworks=list(weighted_subnetworks), bias=bias, logits=logits,
        subnetworks=list(subnetworks or []), complexity_regularization=complexity_regularization)


class MixtureWeightType(object):
  """SCALAR: rank-0 weight; VECTOR: rank-1; MATRIX: rank-2 (weighted.py:139-147)."""
  SCALAR = "scalar"
  VECTOR = "vector"
  MATRIX = "matrix"


def _as_float(x) -> float:
  return float(np.asarray(x, dtype=np.float32))


class ComplexityRegularizedEnsembler(Ensembler):
  """See module docstring; constructor arguments as weighted.py:228-251."""

  def __init__(self, optimizer=None, mixture_weight_type=MixtureWeightType.SCALAR, mixture_weight_initializer=None,
               warm_start_mixture_weights=False, model_dir=None, adanet_lambda=0., adanet_beta=0., use_bias=False,
               name=None):
    if warm_start_mixture_weights:
      if model_dir is None:
        raise ValueError("model_dir cannot be None when warm_start_mixture_weights is True.")
    if mixture_weight_type not in (MixtureWeightType.SCALAR, MixtureWeightType.VECTOR, MixtureWeightType.MATRIX):
      raise ValueError("unknown mixture_weight_type %r" % (mixture_weight_type,))
    self._optimizer = optimizer
    self._mixture_weight_type = mixture_weight_type
    self._mixture_weight_initializer = mixture_weight_initializer
    self._warm_start_mixture_weights = warm_start_mixture_weights
    self._model_dir = model_dir
    self._adanet_lambda = adanet_lambda
    self._adanet_beta = adanet_beta
    self._use_bias = use_bias
    self._name = name

  @property
  def name(self):
    return self._name if self._name else "complexity_regularized"

  # ---- configuration read by the Estimator when lowering to an engine plan ----
  @property
  def mixture_weight_type(self):
    return self._mixture_weight_type

  @property
  def adanet_lambda(self):
    return self._adanet_lambda

  @property
  def adanet_beta(self):
    return self._adanet_beta

  @property
  def use_bias(self):
    return self._use_bias

  @property
  def warm_start_mixture_weights(self):
    return self._warm_start_mixture_weights

  @property
  def optimizer(self):
    return self._optimizer

  # ---- host-side arithmetic pinned by adanet/ensemble/weighted_test.py:147-567 ----
  def _compute_adanet_gamma(self, complexity) -> float:
    """lambda * r(h) + beta (weighted.py:351-358)."""
    if self._adanet_lambda == 0.:
      return float(self._adanet_beta)
    return float(np.float32(self._adanet_lambda) * np.float32(_as_float(complexity)) + np.float32(self._adanet_beta))

  def initial_mixture_weight(self, num_subnetworks: int, last_layer_size: int, logits_size: int) -> np.ndarray:
    """weighted.py:360-366,419-428: SCALAR/VECTOR -> 1/N, MATRIX -> zeros."""
    if self._mixture_weight_initializer is not None:
      shape = {MixtureWeightType.SCALAR: (), MixtureWeightType.VECTOR: (logits_size,),
               MixtureWeightType.MATRIX: (last_layer_size, logits_size)}[self._mixture_weight_type]
      return np.asarray(self._mixture_weight_initializer(shape), dtype=np.float32).reshape(shape)
    if self._mixture_weight_type == MixtureWeightType.SCALAR:
      return np.array(1. / num_subnetworks, dtype=np.float32)
    if self._mixture_weight_type == MixtureWeightType.VECTOR:
      return np.full((logits_size,), 1. / num_subnetworks, dtype=np.float32)
    return np.zeros((last_layer_size, logits_size), dtype=np.float32)

  def complexity_regularization(self, weights: Sequence, complexities: Sequence) -> float:
    """sum_k gamma_k ||w_k||_1; exactly 0 when lambda == beta == 0 (weighted.py:563-604)."""
    if self._adanet_lambda == 0. and self._adanet_beta == 0.:
      return 0.
    total = np.float32(0.)
    for w, c in zip(weights, complexities):
      l1 = np.float32(np.abs(np.asarray(w, dtype=np.float32)).sum(dtype=np.float32))
      total = np.float32(total + np.float32(self._compute_adanet_gamma(c)) * l1)
    return float(total)

  @staticmethod
  def mixture_weight_norms(weights: Sequence):
    """The `mixture_weight_norms/...` and `mixture_weight_fractions/...` summaries (weighted.py:581-594)."""
    norms = [float(np.abs(np.asarray(w, dtype=np.float32)).sum(dtype=np.float32)) for w in weights]
    tot = sum(norms)
    return norms, [n / tot for n in norms]

  # ---- Ensembler interface ----
  def build_ensemble(self, subnetworks, previous_ensemble_subnetworks, features, labels, logits_dimension, training,
                     iteration_step, summary, previous_ensemble, previous_iteration_checkpoint=None):
    """weighted.py:253-336.  Weights start at their default (or warm-started) values;
    the returned ensemble carries values + symbolic member logits for the engine."""
    weighted_subnetworks: List[WeightedSubnetwork] = []
    num_subnetworks = len(subnetworks)
    kept = []
    if previous_ensemble_subnetworks and previous_ensemble:
      num_subnetworks += len(previous_ensemble_subnetworks)
      for ws in previous_ensemble.weighted_subnetworks:
        if not any(ws.subnetwork is s for s in previous_ensemble_subnetworks):
          continue
        kept.append(ws)

    def dims(s):
      last = s.last_layer.shape[-1] if hasattr(s.last_layer, "shape") else None
      lg = s.logits.shape[-1] if hasattr(s.logits, "shape") else logits_dimension
      return last, lg

    for ws in kept:
      last, lg = dims(ws.subnetwork)
      init = None
      if self._warm_start_mixture_weights and ws.weight is not None:
        init = np.asarray(ws.weight, dtype=np.float32)     # learned value of iteration t-1 (weighted.py:275-285)
      w = init if init is not None else self.initial_mixture_weight(num_subnetworks, last, lg)
      weighted_subnetworks.append(WeightedSubnetwork(name=ws.name, iteration_number=ws.iteration_number, weight=w,
                                                     logits=ws.subnetwork.logits, subnetwork=ws.subnetwork))
    for s in subnetworks:
      last, lg = dims(s)
      weighted_subnetworks.append(WeightedSubnetwork(weight=self.initial_mixture_weight(num_subnetworks, last, lg),
                                                     logits=s.logits, subnetwork=s))
    lg0 = dims(weighted_subnetworks[0].subnetwork)[1] or logits_dimension or 1
    bias = np.zeros((lg0,), dtype=np.float32)
    if (previous_ensemble is not None and self._warm_start_mixture_weights and
        len(previous_ensemble.subnetworks) == len(previous_ensemble_subnetworks or []) and
        previous_ensemble.bias is not None):
      bias = np.asarray(previous_ensemble.bias, dtype=np.float32).copy()
    reg = self.complexity_regularization([ws.weight for ws in weighted_subnetworks],
                                         [ws.subnetwork.complexity for ws in weighted_subnetworks])
    return ComplexityRegularized(weighted_subnetworks=weighted_subnetworks, bias=bias,
                                 subnetworks=[ws.subnetwork for ws in weighted_subnetworks],
                                 logits=("weighted_sum", [ws.logits for ws in weighted_subnetworks]),
                                 complexity_regularization=reg)

  def build_train_op(self, ensemble, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    """weighted.py:606-617: no_op without an optimizer; otherwise minimise
    `loss + ensemble.complexity_regularization` -- `loss` is already the
    regularised adanet_loss (ensemble_builder.py:542), so the regulariser is
    counted twice in the optimised objective; the engine reproduces that with
    reg_multiplier = 2."""
    optimizer = self._optimizer
    if callable(optimizer) and not isinstance(optimizer, train.Optimizer):
      optimizer = optimizer()
    if optimizer is None:
      return train.no_op()
    return train.TrainOp("minimize", train.optimizer_from(optimizer), ("sum", loss, "complexity_regularization"),
                         var_list)
