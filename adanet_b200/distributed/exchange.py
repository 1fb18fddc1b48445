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
  mine = torch.full((slots,), float("nan"), dtype=torch.float32, device=device)
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
  flat = torch.cat([t.reshape(-1) for t in tensors]) if rank() == src else \
      torch.empty((sum(t.numel() for t in tensors),), dtype=tensors[0].dtype, device=tensors[0].device)
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
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t.item())
