"""Subnetwork plugin interface: Subnetwork, TrainOpSpec, Builder, Generator.

Mirror of adanet/subnetwork/generator.py (Subnetwork :62-158, TrainOpSpec
:39-59, Builder :161-270, Generator :273-319, SimpleGenerator :323-339): same
names, argument meaning and error behaviour.  Tensors are the symbolic handles
of adanet_b200.graph instead of TF graph tensors.
"""

from __future__ import annotations

import abc
import collections


def _validate_nested_persisted_tensors(persisted_tensors):
  """Raises a ValueError when a nested dict is empty (generator.py:30-36)."""
  for key, entry in persisted_tensors.items():
    if not isinstance(entry, dict):
      continue
    if not entry:
      raise ValueError("Got empty nested dictionary for key: '{}'".format(key))
    _validate_nested_persisted_tensors(entry)


class TrainOpSpec(collections.namedtuple("TrainOpSpec", ["train_op", "chief_hooks", "hooks"])):
  """A train op plus (ignored: no session) hooks; generator.py:39-59."""

  def __new__(cls, train_op, chief_hooks=None, hooks=None):
    return super(TrainOpSpec, cls).__new__(cls, train_op, tuple(chief_hooks) if chief_hooks else (),
                                           tuple(hooks) if hooks else ())


class Subnetwork(collections.namedtuple(
    "Subnetwork", ["last_layer", "logits", "complexity", "persisted_tensors", "shared", "local_init_ops"])):
  """An AdaNet subnetwork *h* (generator.py:62-158).

  Raises ValueError when last_layer / logits / complexity is None, when only
  one of logits / last_layer is a dict, or when persisted_tensors is malformed.
  """

  def __new__(cls, last_layer, logits, complexity, persisted_tensors=None, shared=None, local_init_ops=None):
    if last_layer is None:
      raise ValueError("last_layer not provided")
    if logits is None:
      raise ValueError("logits not provided")
    if isinstance(logits, dict) and not isinstance(last_layer, dict):
      raise ValueError("if logits is a dict last_layer must also be a dict")
    if isinstance(last_layer, dict) and not isinstance(logits, dict):
      raise ValueError("if last_layer is a dict logits must also be a dict")
    if complexity is None:
      raise ValueError("complexity not provided")
    if persisted_tensors is not None:
      if not isinstance(persisted_tensors, dict):
        raise ValueError("persisted_tensors must be a dict")
      _validate_nested_persisted_tensors(persisted_tensors)
    local_init_ops = tuple(local_init_ops) if local_init_ops else ()
    return super(Subnetwork, cls).__new__(cls, last_layer=last_layer, logits=logits, complexity=complexity,
                                          persisted_tensors=persisted_tensors, shared=shared,
                                          local_init_ops=local_init_ops)


class Builder(abc.ABC):
  """Interface for a subnetwork builder (generator.py:161-270)."""

  @property
  @abc.abstractmethod
  def name(self):
    """Unique name of the subnetwork within an iteration."""

  @abc.abstractmethod
  def build_subnetwork(self, features, labels, logits_dimension, training, iteration_step, summary,
                       previous_ensemble=None):
    """Returns the candidate `Subnetwork`.  `labels` and `config` are optional
    arguments detected by name (adanet/core/ensemble_builder.py:737-746)."""

  @abc.abstractmethod
  def build_subnetwork_train_op(self, subnetwork, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    """Returns the op (or TrainOpSpec) that trains `var_list` on `loss`."""

  def build_subnetwork_report(self):
    """Optional `adanet.subnetwork.Report` (generator.py:259-270)."""
    return None


class Generator(abc.ABC):
  """Interface for a candidate subnetwork generator (generator.py:273-319)."""

  @abc.abstractmethod
  def generate_candidates(self, previous_ensemble, iteration_number, previous_ensemble_reports, all_reports):
    """Returns the list of `Builder`s to train this iteration.  Must be
    deterministic for fixed arguments (it is called on every rank)."""


class SimpleGenerator(Generator):
  """Always generates the given list of builders (generator.py:323-339)."""

  def __init__(self, subnetwork_builders):
    self._subnetwork_builders = subnetwork_builders

  def generate_candidates(self, previous_ensemble, iteration_number, previous_ensemble_reports, all_reports):
    return self._subnetwork_builders
