"""adanet.replay mirror (adanet/replay/__init__.py:28-59): deterministic replay of a
previous search by overriding the per-iteration best ensemble index
(used at adanet/core/estimator.py:1433-1438)."""

__all__ = ["Config"]


class Config(object):
  """Defines how to deterministically replay an AdaNet model search."""

  def __init__(self, best_ensemble_indices=None):
    self._best_ensemble_indices = best_ensemble_indices

  @property
  def best_ensemble_indices(self):
    """The best ensemble indices per iteration."""
    return self._best_ensemble_indices

  def get_best_ensemble_index(self, iteration_number):
    """Returns the best ensemble index given an iteration number, else None."""
    if self._best_ensemble_indices and iteration_number < len(self._best_ensemble_indices):
      return self._best_ensemble_indices[iteration_number]
    return None
