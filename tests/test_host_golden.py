"""Host-side mirror of the reference interface vs fixtures produced by executing
the reference's own modules (tests/golden/make_golden.py).  CPU only."""

import json
import os
import re

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
  with open(os.path.join(GOLD, name)) as f:
    return json.load(f)


def test_round_robin_truth_tables():
  # adanet/distributed/placement_test.py:66-298 (truth tables), generated from the reference class itself
  from adanet_b200.distributed import ClusterConfig, RoundRobinStrategy
  rows = _load("placement.json")["round_robin"]
  assert len(rows) > 300
  for row in rows:
    s = RoundRobinStrategy(drop_remainder=row["drop_remainder"])
    s.config = ClusterConfig(row["num_workers"], row["worker_index"])
    ns = row["num_subnetworks"]
    assert s.should_build_ensemble(ns) == row["build_ensemble"], row
    assert [s.should_build_subnetwork(ns, i) for i in range(ns)] == row["build_subnetwork"], row
    assert s.should_train_subnetworks(ns) == row["train_subnetworks"], row


def test_replication_strategy():
  from adanet_b200.distributed import ClusterConfig, ReplicationStrategy
  want = _load("placement.json")["replication"]
  s = ReplicationStrategy()
  s.config = ClusterConfig(3, 1)
  assert s.should_build_ensemble(3) == want["build_ensemble"]
  assert [s.should_build_subnetwork(3, i) for i in range(3)] == want["build_subnetwork"]
  assert s.should_train_subnetworks(3) == want["train_subnetworks"]


def test_reference_round_robin_8_workers_8_subnetworks_finding():
  # SURVEY.md section 0 finding 2: worker0 -> ensemble only, worker1 -> subnets [0,7], workers 2-7 -> [1]..[6]
  from adanet_b200.distributed import ClusterConfig, RoundRobinStrategy
  got = []
  for wi in range(8):
    s = RoundRobinStrategy()
    s.config = ClusterConfig(8, wi)
    got.append([i for i in range(8) if s.should_build_subnetwork(8, i)] if s.should_train_subnetworks(8) else "ens")
  assert got == ["ens", [0, 7], [1], [2], [3], [4], [5], [6]]


def test_colocated_strategy_is_i_mod_g():
  from adanet_b200.distributed import ClusterConfig, ColocatedStrategy
  for g in (1, 2, 4, 8):
    seen = []
    for r in range(g):
      s = ColocatedStrategy()
      s.config = ClusterConfig(g, r)
      own = s.owned(32)
      assert all(i % g == r for i in own)
      assert s.should_train_subnetworks(32) and s.should_build_ensemble(32)
      seen += own
    assert sorted(seen) == list(range(32))


def test_cabi_header_symbols_are_exported(built_lib):
  """The shared library loads and exports every entry point include/adanet_b200.h declares."""
  from adanet_b200 import _lib
  hdr = open(os.path.join(ROOT, "include", "adanet_b200.h")).read()
  declared = sorted(set(re.findall(r"\b(adn_[a-z0-9_]+)\s*\(", hdr)))
  assert len(declared) >= 12
  for name in declared:
    assert hasattr(built_lib, name), "missing export %s" % name
  assert sorted(_lib.EXPORTS) == declared
  assert _lib.query(_lib.Q_VERSION) == 100


def test_cabi_argument_validation_without_gpu(built_lib):
  """Error behaviour of the boundary: bad arguments fail loudly with a message, before any launch."""
  from adanet_b200 import _lib
  rc = built_lib.adn_dense_fwd(None, None, None, None, 4, 4, 4, 0, None, 0, None)
  assert rc == -22 and b"null" in built_lib.adn_last_error()
  rc = built_lib.adn_dense_fwd(8, 8, None, 8, 0, 4, 4, 0, None, 0, None)
  assert rc == -22
  rc = built_lib.adn_dense_fwd(8, 8, None, 8, 4, 4, 4, 7, None, 0, None)
  assert rc == -22 and b"act" in built_lib.adn_last_error()
  rc = built_lib.adn_ema_update(None, None, 0.9, None)
  assert rc == -22
  with pytest.raises(_lib.AdnError):
    _lib.check(built_lib.adn_set_dense_path(9), "adn_set_dense_path")
  assert _lib.query(_lib.Q_DENSE_BWD_WS, 4096, 100, 1024) > 0
  assert _lib.query(_lib.Q_HEAD_WS, 4096, 10, 5) > 0


def test_engine_fails_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  from adanet_b200 import _lib
  from adanet_b200.core import engine as eng
  with pytest.raises(_lib.AdnError):
    eng.IterationPlan(0, [], [], eng.EnsemblerPlanSpec(), 8, 4, 2)
