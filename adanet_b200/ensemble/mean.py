"""MeanEnsembler (mirror of adanet/ensemble/mean.py:29-135): uniform mean of the
*new* subnetworks' logits, no trainable ensemble variables."""

from __future__ import annotations

import collections

from adanet_b200 import train
from adanet_b200.ensemble.ensembler import Ensemble
from adanet_b200.ensemble.ensembler import Ensembler


class MeanEnsemble(collections.namedtuple("MeanEnsemble", ["logits", "subnetworks", "predictions"]), Ensemble):
  """Mean ensemble (mean.py:29-53)."""

  MEAN_LAST_LAYER = "mean_last_layer"

  def __new__(cls, logits, subnetworks=None, predictions=None):
    return super(MeanEnsemble, cls).__new__(cls, logits=logits, subnetworks=list(subnetworks or []),
                                            predictions=predictions)


class MeanEnsembler(Ensembler):
  """Takes the mean of the logits of its subnetworks (mean.py:56-135)."""

  def __init__(self, name=None, add_mean_last_layer_predictions=False):
    self._name = name
    self._add_mean_last_layer_predictions = add_mean_last_layer_predictions

  @property
  def name(self):
    return self._name if self._name else "mean"

  def build_ensemble(self, subnetworks, previous_ensemble_subnetworks, features, labels, logits_dimension, training,
                     iteration_step, summary, previous_ensemble, previous_iteration_checkpoint=None):
    if self._add_mean_last_layer_predictions:
      shapes = [tuple(s.last_layer.shape) for s in subnetworks]
      if any(sh != shapes[0] for sh in shapes):
        raise ValueError("Shape of `last_layer` tensors must be same if setting "
                         "`add_mean_last_layer_predictions` to True. Found %s vs %s." % (shapes[0], shapes[-1]))
    # symbolic: ("mean", [logits...]); the engine realises it as SCALAR weights 1/N over the new members
    return MeanEnsemble(logits=("mean", [s.logits for s in subnetworks]), subnetworks=subnetworks, predictions=None)

  def build_train_op(self, ensemble, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    return train.no_op()
