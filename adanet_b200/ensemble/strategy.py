"""Search strategies for ensemble candidates (mirror of adanet/ensemble/strategy.py:26-117)."""

from __future__ import annotations

import abc
import collections


class Candidate(collections.namedtuple("Candidate", ["name", "subnetwork_builders",
                                                     "previous_ensemble_subnetwork_builders"])):
  """An ensemble candidate found during the search phase (strategy.py:26-49)."""

  def __new__(cls, name, subnetwork_builders, previous_ensemble_subnetwork_builders):
    return super(Candidate, cls).__new__(cls, name=name, subnetwork_builders=tuple(subnetwork_builders),
                                         previous_ensemble_subnetwork_builders=tuple(
                                             previous_ensemble_subnetwork_builders or []))


class Strategy(abc.ABC):
  """An abstract ensemble strategy (strategy.py:52-76)."""

  @abc.abstractmethod
  def generate_ensemble_candidates(self, subnetwork_builders, previous_ensemble_subnetwork_builders):
    """Returns the `Candidate`s to train and consider this iteration."""


class SoloStrategy(Strategy):
  """An ensemble of one: prunes every previous subnetwork (strategy.py:79-94)."""

  def generate_ensemble_candidates(self, subnetwork_builders, previous_ensemble_subnetwork_builders):
    return [Candidate("{}_solo".format(b.name), [b], None) for b in subnetwork_builders]


class GrowStrategy(Strategy):
  """Greedily grows the ensemble one subnetwork at a time (strategy.py:97-106)."""

  def generate_ensemble_candidates(self, subnetwork_builders, previous_ensemble_subnetwork_builders):
    return [Candidate("{}_grow".format(b.name), [b], previous_ensemble_subnetwork_builders)
            for b in subnetwork_builders]


class AllStrategy(Strategy):
  """Ensembles all subnetworks of the iteration (strategy.py:109-117)."""

  def generate_ensemble_candidates(self, subnetwork_builders, previous_ensemble_subnetwork_builders):
    return [Candidate("all", subnetwork_builders, previous_ensemble_subnetwork_builders)]
