"""AutoEnsembleEstimator (mirror of adanet/autoensemble/estimator.py:28-220): an
adanet Estimator whose generator is a fixed pool of sub-estimators."""

from __future__ import annotations

from adanet_b200.autoensemble.common import _GeneratorFromCandidatePool
from adanet_b200.core.estimator import Estimator


class AutoEnsembleEstimator(Estimator):
  """Learns to ensemble the models of `candidate_pool` (list, dict name->estimator, or a
  callable(config[, iteration_number]) returning either)."""

  def __init__(self, head, candidate_pool, max_iteration_steps, ensemblers=None, ensemble_strategies=None,
               logits_fn=None, last_layer_fn=None, evaluator=None, metric_fn=None, force_grow=False,
               adanet_loss_decay=.9, worker_wait_timeout_secs=7200, model_dir=None, config=None, **kwargs):
    subnetwork_generator = _GeneratorFromCandidatePool(candidate_pool, logits_fn, last_layer_fn)
    super(AutoEnsembleEstimator, self).__init__(
        head=head, subnetwork_generator=subnetwork_generator, max_iteration_steps=max_iteration_steps,
        ensemblers=ensemblers, ensemble_strategies=ensemble_strategies, evaluator=evaluator, metric_fn=metric_fn,
        force_grow=force_grow, adanet_loss_decay=adanet_loss_decay, worker_wait_timeout_secs=worker_wait_timeout_secs,
        model_dir=model_dir, config=config, **kwargs)
