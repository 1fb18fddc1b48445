"""Ensembler plugin interface (mirror of adanet/ensemble/ensembler.py:26-150)."""

from __future__ import annotations

import abc
import collections


class TrainOpSpec(collections.namedtuple("TrainOpSpec", ["train_op", "chief_hooks", "hooks"])):
  """Ensembler training op + (ignored) hooks; ensembler.py:26-47."""

  def __new__(cls, train_op, chief_hooks=None, hooks=None):
    return super(TrainOpSpec, cls).__new__(cls, train_op, tuple(chief_hooks) if chief_hooks else (),
                                           tuple(hooks) if hooks else ())


class Ensemble(abc.ABC):
  """An abstract ensemble of subnetworks (ensembler.py:50-70)."""

  @property
  @abc.abstractmethod
  def logits(self):
    """Ensemble logits."""

  @property
  @abc.abstractmethod
  def subnetworks(self):
    """Ordered iterable of the ensemble's subnetworks."""

  @property
  def predictions(self):
    return None


class Ensembler(abc.ABC):
  """An abstract ensembler (ensembler.py:73-150)."""

  @property
  @abc.abstractmethod
  def name(self):
    """This ensembler's unique string name."""

  @abc.abstractmethod
  def build_ensemble(self, subnetworks, previous_ensemble_subnetworks, features, labels, logits_dimension, training,
                     iteration_step, summary, previous_ensemble, previous_iteration_checkpoint):
    """Builds an ensemble of subnetworks; returns an `Ensemble`."""

  @abc.abstractmethod
  def build_train_op(self, ensemble, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    """Returns a train op (or TrainOpSpec) for the ensemble's own variables."""
