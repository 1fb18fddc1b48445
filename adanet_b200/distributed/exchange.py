"""End-of-iteration exchange between GPU ranks.

The reference synchronises workers through parameter servers and a shared
`model_dir` (adanet/core/estimator.py:951-984, adanet/core/iteration.py:81-105)
and has no collective at all.  On one NVSwitch box the only data that must
cross GPUs is, once per AdaNet iteration:

  1. the K candidates' EMA adanet losses  -> all_gather (K fp32) so every rank
     computes the same nanargmin (adanet/core/estimator.py:1491-1495);
  2. the winner's new parameters           -> broadcast from its owner rank,
     replacing the reference's shared checkpoint (estimator.py:1357-1406).

Both are microsecond-scale over NVLink; there is no data-path collective
inside a training step.  The functions take plain torch tensors and a process
group so the same code runs on NCCL (GPU) and gloo (CPU tests).
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def world() -> int:
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _comm_device(device):
  """Where collective payloads live: the GPU for NCCL; host memory for gloo (CPU tests, and several ranks sharing
  one GPU when a box has fewer GPUs than ranks -- the payloads are a few floats or a few MB once per iteration)."""
  if dist.is_available() and dist.is_initialized() and dist.get_backend() == "gloo":
    return torch.device("cpu")
  return device


def owner_of(candidate_index: int, world_size: int) -> int:
  """Colocated placement: candidate i lives on rank i % G."""
  return candidate_index % world_size


def round_robin_owners(num_candidates: int, world_size: int) -> List[int]:
  return [owner_of(i, world_size) for i in range(num_candidates)]


def balanced_owners(costs: Sequence[float], world_size: int) -> List[int]:
  """Cost-balanced colocated placement (longest-processing-time first): the heaviest candidate goes to the
  least-loaded rank (ties: lowest rank).  Candidates differ by >100x in FLOPs across a width sweep, so
  `i % G` leaves most GPUs waiting for the rank that drew the widest ones; every rank computes this same
  mapping from the candidate specs, no communication."""
  order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
  load = [0.0] * world_size
  owners = [0] * len(costs)
  for i in order:
    r = min(range(world_size), key=lambda q: (load[q], q))
    owners[i] = r
    load[r] += float(costs[i])
  return owners


def component_owners(costs: Sequence[float], groups: Sequence[Sequence[int]], world_size: int,
                     placement: str = "balanced") -> List[int]:
  """Owner rank of every subnetwork when `groups[j]` lists the subnetworks candidate ensemble j reads.

  Subnetworks read by one ensemble must share a GPU, so the connected components of the "shares an ensemble"
  relation are placed as units (cost = sum of member costs): LPT for "balanced", component-index % G for
  "round_robin".  With one subnetwork per ensemble (GrowStrategy) this is exactly balanced_owners / i % G."""
  n = len(costs)
  comp = list(range(n))

  def find(i):
    while comp[i] != i:
      comp[i] = comp[comp[i]]
      i = comp[i]
    return i

  for grp in groups:
    for b in grp[1:]:
      ra, rb = find(grp[0]), find(b)
      comp[max(ra, rb)] = min(ra, rb)
  roots = sorted({find(i) for i in range(n)})
  comp_cost = [sum(float(costs[i]) for i in range(n) if find(i) == q) for q in roots]
  comp_owner = balanced_owners(comp_cost, world_size) if placement == "balanced" else \
      round_robin_owners(len(roots), world_size)
  return [comp_owner[roots.index(find(i))] for i in range(n)]


def sharded_placement(costs: Sequence[float], world_size: int, batch: int, tolerance: float = 1.15) -> List[List[int]]:
  """Row-sharded placement: candidate i is trained by `ranks[i]` (a power-of-two number of ranks, increasing), shard j
  of its minibatch rows on ranks[i][j].

  Whole-candidate placement cannot balance a width sweep: in BASELINE configs[2] the H=1024 candidate alone is 46 % of
  the step, so 8 GPUs could never be more than 2.2x faster than one (round-1 SCALE: efficiency 0.31).  A candidate
  whose cost exceeds `tolerance` x the mean rank load is therefore split by ROWS over 2, 4, ... ranks: every rank
  runs the same forward / backward on its slice of the minibatch and the slices' gradients are averaged across the
  candidate's ranks once per step (core/engine.ShardComm) -- data parallelism inside a candidate, candidate
  parallelism across them.  The pieces are then placed longest-first on the least-loaded rank that does not already
  hold a piece of the same candidate.  Every rank computes the same mapping from the specs, no communication."""
  n = len(costs)
  total = float(sum(costs))
  target = total / max(1, world_size)
  shards = []
  for c in costs:
    g = 1
    while float(c) / g > tolerance * target and g * 2 <= world_size and batch % (g * 2) == 0:
      g *= 2
    shards.append(g)
  pieces = sorted(((float(costs[i]) / shards[i], i, j) for i in range(n) for j in range(shards[i])),
                  key=lambda p: (-p[0], p[1], p[2]))
  load = [0.0] * world_size
  ranks: List[List[int]] = [[] for _ in range(n)]
  for cost, i, _ in pieces:
    r = min((q for q in range(world_size) if q not in ranks[i]), key=lambda q: (load[q], q))
    ranks[i].append(r)
    load[r] += cost
  return [sorted(r) for r in ranks]


_GROUPS = {}


def subgroup(ranks: Sequence[int]):
  """Process group of `ranks` (cached).  torch.distributed.new_group is collective over the WHOLE job: every rank must
  call this for every group in the same order, member or not (AdaNetSearch.build_iteration does)."""
  key = tuple(int(r) for r in ranks)
  if len(key) == world():
    return None      # the default group
  if key not in _GROUPS:
    _GROUPS[key] = dist.new_group(list(key))
  return _GROUPS[key]


def owned_indices(num_candidates: int, rank_: int, world_size: int, owners: Optional[Sequence[int]] = None) -> List[int]:
  owners = owners if owners is not None else round_robin_owners(num_candidates, world_size)
  return [i for i in range(num_candidates) if owners[i] == rank_]


def gather_candidate_losses(local_losses: Sequence[float], num_candidates: int, device=None, group=None,
                            owners: Optional[Sequence[int]] = None) -> List[float]:
  """all_gather of the per-candidate EMA losses, returned in candidate order.

  local_losses[j] belongs to the j-th candidate this rank owns (increasing candidate index).  Every rank
  pads to the largest per-rank count with NaN, gathers, and un-interleaves with the shared owner map.
  """
  g, r = world(), rank()
  if g == 1:
    return [float(v) for v in local_losses]
  owners = list(owners) if owners is not None else round_robin_owners(num_candidates, g)
  slots = max(1, max(owners.count(q) for q in range(g)))
  mine = torch.full((slots,), float("nan"), dtype=torch.float32, device=_comm_device(device))
  for j, v in enumerate(local_losses):
    mine[j] = float(v)
  parts = [torch.empty_like(mine) for _ in range(g)]
  dist.all_gather(parts, mine, group=group)   # ncclAllGather on GPU, gloo on CPU
  out = torch.stack(parts).cpu()
  pos, seen = [], [0] * g
  for i in range(num_candidates):
    pos.append(seen[owners[i]])
    seen[owners[i]] += 1
  return [float(out[owners[i], pos[i]]) for i in range(num_candidates)]


def broadcast_tensors(tensors: Sequence[torch.Tensor], src: int, group=None) -> None:
  """Broadcast the winner's parameters from its owner; one flat message."""
  if world() == 1:
    return
  cdev = _comm_device(tensors[0].device)
  flat = torch.cat([t.reshape(-1) for t in tensors]).to(cdev) if rank() == src else \
      torch.empty((sum(t.numel() for t in tensors),), dtype=tensors[0].dtype, device=cdev)
  dist.broadcast(flat, src=src, group=group)
  if rank() != src:
    off = 0
    for t in tensors:
      n = t.numel()
      t.copy_(flat[off:off + n].view_as(t))
      off += n


def max_over_ranks(value: float, device=None, group=None) -> float:
  """Timing reduction used by bench.py (max over ranks of a device-measured time)."""
  if world() == 1:
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device(device))
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t.item())
