"""adanet.distributed mirror (adanet/distributed/__init__.py:26-34)."""

from adanet_b200.distributed.placement import ClusterConfig
from adanet_b200.distributed.placement import ColocatedStrategy
from adanet_b200.distributed.placement import PlacementStrategy
from adanet_b200.distributed.placement import ReplicationStrategy
from adanet_b200.distributed.placement import RoundRobinStrategy

__all__ = ["PlacementStrategy", "ReplicationStrategy", "RoundRobinStrategy", "ColocatedStrategy", "ClusterConfig"]
