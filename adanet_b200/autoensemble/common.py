"""Candidate-pool plumbing of AutoEnsembleEstimator.

Mirror of adanet/autoensemble/common.py: `AutoEnsembleSubestimator` (:63-93),
`_BuilderFromSubestimator` (:96-198: logits AND last_layer both come from the
sub-estimator's logits, complexity 0, the sub-estimator's own train op) and
`_GeneratorFromCandidatePool` (:218-268: dict pools sorted by name, list pools
named "{ClassName}{index}", callable pools called with config[/iteration_number]).
"""

from __future__ import annotations

import collections
import inspect

from adanet_b200 import estimators
from adanet_b200 import subnetwork as subnetwork_lib


class AutoEnsembleSubestimator(collections.namedtuple("AutoEnsembleSubestimator",
                                                      ["estimator", "train_input_fn", "prediction_only"])):
  """A sub-estimator with optional bagging input (common.py:63-93): with `train_input_fn` the subnetwork trains on
  minibatches of that input_fn (same `(features, labels)` batch conventions and batch size as the Estimator's
  `input_fn`), one step before each main step, and the ensembles read its forward on the shared minibatch."""

  def __new__(cls, estimator, train_input_fn=None, prediction_only=False):
    return super(AutoEnsembleSubestimator, cls).__new__(cls, estimator, train_input_fn, prediction_only)


def _convert_to_subestimator(candidate):
  """common.py:200-215."""
  if isinstance(candidate, AutoEnsembleSubestimator):
    return lambda config: candidate
  if isinstance(candidate, estimators.SubEstimator):
    return lambda config: AutoEnsembleSubestimator(candidate)
  if callable(candidate):
    return candidate
  raise ValueError("subestimator in candidate_pool must have type adanet_b200.estimators.SubEstimator or "
                   "adanet.AutoEnsembleSubestimator but got {}".format(candidate.__class__))


class _BuilderFromSubestimator(subnetwork_lib.Builder):
  """An adanet Builder from a sub-estimator (common.py:96-198)."""

  def __init__(self, name, subestimator, logits_fn, last_layer_fn, config):
    self._name = name
    self._subestimator = subestimator(config)
    self._logits_fn = logits_fn
    self._last_layer_fn = last_layer_fn

  @property
  def name(self):
    return self._name

  def build_subnetwork(self, features, labels, logits_dimension, training, iteration_step, summary,
                       previous_ensemble=None, config=None):
    sub = self._subestimator
    # bagging (common.py:151-180): the engine is told to train this subnetwork on its own input_fn
    self.bagging_train_input_fn = sub.train_input_fn if (training and not sub.prediction_only) else None
    logits = sub.estimator.build_logits(features, logits_dimension)
    if self._logits_fn is not None:
      logits = self._logits_fn(logits)
    last_layer = self._last_layer_fn(logits) if self._last_layer_fn else logits   # common.py:115-118
    return subnetwork_lib.Subnetwork(last_layer=last_layer, logits=logits, complexity=0., shared=None)   # :188

  def build_subnetwork_train_op(self, subnetwork, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    if self._subestimator.prediction_only:
      from adanet_b200 import train
      return train.no_op()
    return self._subestimator.estimator.train_op(loss, var_list)   # the sub-estimator's own train op (:196-198)


class _GeneratorFromCandidatePool(subnetwork_lib.Generator):
  """An adanet Generator from a pool of sub-estimators (common.py:218-268)."""

  def __init__(self, candidate_pool, logits_fn, last_layer_fn):
    self._candidate_pool = candidate_pool
    self._logits_fn = logits_fn
    self._last_layer_fn = last_layer_fn

  def generate_candidates(self, previous_ensemble, iteration_number, previous_ensemble_reports, all_reports, config):
    assert config
    builders = []
    pool = self._maybe_call_candidate_pool(config, iteration_number)
    if isinstance(pool, dict):
      for name in sorted(pool):
        builders.append(_BuilderFromSubestimator(name, _convert_to_subestimator(pool[name]), self._logits_fn,
                                                 self._last_layer_fn, config))
      return builders
    for i, est in enumerate(pool):
      inner = est.estimator if isinstance(est, AutoEnsembleSubestimator) else est
      name = "{class_name}{index}".format(class_name=inner.__class__.__name__, index=i)
      builders.append(_BuilderFromSubestimator(name, _convert_to_subestimator(est), self._logits_fn,
                                               self._last_layer_fn, config))
    return builders

  def _maybe_call_candidate_pool(self, config, iteration_number):
    if callable(self._candidate_pool):
      args = inspect.signature(self._candidate_pool).parameters
      if "iteration_number" in args:
        return self._candidate_pool(config=config, iteration_number=iteration_number)
      return self._candidate_pool(config=config)
    return self._candidate_pool
