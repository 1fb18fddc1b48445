"""Heads: the loss the engine's fused head kernels compute.

Stand-ins for the tf.estimator heads the reference is constructed with
(`head.create_estimator_spec(...)` at adanet/core/ensemble_builder.py:571-583):
MultiClassHead = mean sparse softmax cross-entropy, RegressionHead = mean
squared error, BinaryClassHead = mean sigmoid cross-entropy (all
SUM_OVER_BATCH_SIZE, the v2-head default the reference's tests use,
adanet/core/testing_utils.py:236-239).
"""

from __future__ import annotations

from adanet_b200 import graph


class Head:
  loss_kind = None
  logits_dimension = None
  name = None

  def create_loss(self, logits):
    return graph.Loss(logits, self.loss_kind)


class MultiClassHead(Head):
  loss_kind = "softmax_xent"

  def __init__(self, n_classes: int, name=None):
    if n_classes is None or n_classes < 2:
      raise ValueError("n_classes must be >= 2")
    self.n_classes, self.logits_dimension, self.name = n_classes, n_classes, name


class RegressionHead(Head):
  loss_kind = "mse"

  def __init__(self, label_dimension: int = 1, name=None):
    self.logits_dimension, self.name = label_dimension, name


class BinaryClassHead(Head):
  loss_kind = "sigmoid_xent"

  def __init__(self, name=None):
    self.logits_dimension, self.name = 1, name
