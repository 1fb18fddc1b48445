"""Canned sub-estimators for AutoEnsembleEstimator candidate pools.

Stand-ins for tf.estimator.LinearEstimator / DNNEstimator (whose arithmetic is
TensorFlow's): a sub-estimator here is just the recipe
its `model_fn` would build -- a dense chain over the feature columns and the
optimizer that trains it -- expressed in the adanet_b200.graph vocabulary so the
B200 engine can run it.
"""

from __future__ import annotations

from typing import Optional, Sequence

from adanet_b200 import graph
from adanet_b200 import train


class SubEstimator:
  """Base: build_logits(features, logits_dimension) -> (logits, last_layer); optimizer."""

  def __init__(self, feature_columns, optimizer, seed: Optional[int] = None, kernel_initializers=None):
    if not feature_columns:
      raise ValueError("feature_columns must not be empty")
    self.feature_columns, self.optimizer, self.seed = list(feature_columns), optimizer, seed
    self.kernel_initializers = kernel_initializers

  def _init(self, i):
    if self.kernel_initializers is not None:
      return self.kernel_initializers[i]
    return graph.glorot_uniform_initializer(None if self.seed is None else self.seed + i)

  def hidden_units(self) -> Sequence[int]:
    return ()

  def build_logits(self, features, logits_dimension):
    x = graph.input_layer(features, self.feature_columns)
    for i, u in enumerate(self.hidden_units()):
      x = graph.dense(x, u, activation=graph.relu, kernel_initializer=self._init(i), name="hiddenlayer_%d" % i)
    logits = graph.dense(x, logits_dimension, kernel_initializer=self._init(len(self.hidden_units())), name="logits")
    return logits

  def train_op(self, loss, var_list):
    return train.TrainOp("minimize", train.optimizer_from(self.optimizer), loss, var_list)


class LinearEstimator(SubEstimator):
  """logits = input_layer @ W + b (tf.estimator.LinearEstimator analogue)."""


class DNNEstimator(SubEstimator):
  """input_layer -> [dense+relu](hidden_units) -> dense (tf.estimator.DNNEstimator analogue)."""

  def __init__(self, feature_columns, hidden_units, optimizer, seed=None, kernel_initializers=None):
    super().__init__(feature_columns, optimizer, seed, kernel_initializers)
    if not hidden_units:
      raise ValueError("hidden_units must not be empty")
    self._hidden_units = list(hidden_units)

  def hidden_units(self):
    return self._hidden_units
