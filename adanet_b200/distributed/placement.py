"""Placement strategies: which worker (GPU rank) builds / trains what.

API mirror of adanet/distributed/placement.py (PlacementStrategy :31-100,
ReplicationStrategy :103-131, RoundRobinStrategy :134-320).  The reference's
decisions are reproduced from a "task set" formulation (the one its own TODO
at placement.py:171-225 sketches): task 0 = build/train the ensembles, task
k+1 = train subnetwork k.  Truth tables are checked against fixtures generated
by executing the reference class (tests/golden/placement.json).

Parameter-server device functions (placement.py:287-320) are a TF-graph
artefact with no equivalent on a single NVSwitch box and are out of scope
(SURVEY.md section 2 row 8); `subnetwork_devices` is a no-op context here.

`ColocatedStrategy` is the engine's own mapping: under
GrowStrategy each candidate ensemble contains exactly one new subnetwork, so
candidate i -- its subnetwork, ensemble head, mixture weights and EMA -- lives
on GPU `i % num_gpus`, and every GPU trains.
"""

from __future__ import annotations

import abc
import contextlib
from typing import List


class ClusterConfig:
  """The three RunConfig fields the strategies read (placement.py:238-244)."""

  def __init__(self, num_worker_replicas: int = 1, global_id_in_cluster: int = 0, num_ps_replicas: int = 0):
    self.num_worker_replicas = num_worker_replicas
    self.global_id_in_cluster = global_id_in_cluster
    self.num_ps_replicas = num_ps_replicas

  @classmethod
  def from_env(cls):
    import os
    return cls(int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))


class PlacementStrategy(abc.ABC):
  """Abstract placement strategy (placement.py:31-100)."""

  @property
  def config(self):
    return self._config

  @config.setter
  def config(self, config):
    self._config = config

  @abc.abstractmethod
  def should_build_ensemble(self, num_subnetworks: int) -> bool:
    """Whether to build the ensemble on the current worker."""

  @abc.abstractmethod
  def should_build_subnetwork(self, num_subnetworks: int, subnetwork_index: int) -> bool:
    """Whether to build the given subnetwork on the current worker."""

  @abc.abstractmethod
  def should_train_subnetworks(self, num_subnetworks: int) -> bool:
    """Whether to train subnetworks on the current worker."""

  @contextlib.contextmanager
  def subnetwork_devices(self, num_subnetworks: int, subnetwork_index: int):
    """Device context for a subnetwork's ops; a no-op (no parameter servers)."""
    yield


class ReplicationStrategy(PlacementStrategy):
  """Every worker builds and trains everything (placement.py:103-131)."""

  def should_build_ensemble(self, num_subnetworks):
    return True

  def should_build_subnetwork(self, num_subnetworks, subnetwork_index):
    return True

  def should_train_subnetworks(self, num_subnetworks):
    return True


class RoundRobinStrategy(PlacementStrategy):
  """Worker 0 (mod k+1) owns the ensembles; the others own subnetworks round-robin.

  Same decisions as placement.py:228-285 for every (num_workers, worker_index,
  num_subnetworks, drop_remainder).
  """

  def __init__(self, drop_remainder: bool = False, dedicate_parameter_servers: bool = True):
    self._drop_remainder = drop_remainder
    self._dedicate_parameter_servers = dedicate_parameter_servers

  @property
  def _num_workers(self) -> int:
    return self.config.num_worker_replicas

  @property
  def _worker_index(self) -> int:
    return self.config.global_id_in_cluster or 0

  def _task(self, num_subnetworks: int) -> int:
    # one ensemble task + one task per subnetwork, assigned modulo (placement.py:257)
    return self._worker_index % (num_subnetworks + 1)

  def _subnetwork_slots(self, num_subnetworks: int) -> int:
    """How many distinct subnetwork workers exist in this worker's "round"."""
    tasks = num_subnetworks + 1
    full_rounds = self._num_workers // tasks
    remainder = self._num_workers % tasks
    in_partial_round = remainder != 0 and self._worker_index >= full_rounds * tasks
    return remainder - 1 if in_partial_round else num_subnetworks

  def should_build_ensemble(self, num_subnetworks):
    return num_subnetworks == 1 or self._task(num_subnetworks) == 0

  def should_build_subnetwork(self, num_subnetworks, subnetwork_index):
    if num_subnetworks == 1:
      return True
    task = self._task(num_subnetworks)
    if task == 0:
      return True    # the ensemble worker needs every subnetwork's logits
    slot = task - 1
    if self._drop_remainder:
      return slot == subnetwork_index
    return slot == subnetwork_index % self._subnetwork_slots(num_subnetworks)

  def should_train_subnetworks(self, num_subnetworks):
    if num_subnetworks == 1 or self._num_workers == 1:
      return True
    return not self.should_build_ensemble(num_subnetworks)


class ColocatedStrategy(PlacementStrategy):
  """Engine mapping: candidate i (subnetwork + its `*_grow` ensemble) -> GPU i % G."""

  def should_build_ensemble(self, num_subnetworks):
    return True    # every GPU builds the candidate ensembles of the subnetworks it owns

  def should_build_subnetwork(self, num_subnetworks, subnetwork_index):
    return subnetwork_index % self.config.num_worker_replicas == (self.config.global_id_in_cluster or 0)

  def should_train_subnetworks(self, num_subnetworks):
    return True

  def owner(self, subnetwork_index: int) -> int:
    return subnetwork_index % self.config.num_worker_replicas

  def owned(self, num_subnetworks: int) -> List[int]:
    return [i for i in range(num_subnetworks) if self.should_build_subnetwork(num_subnetworks, i)]
