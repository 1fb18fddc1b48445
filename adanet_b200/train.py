"""TF1-semantics optimizers and train ops for builders / ensemblers.

`optimizer.minimize(loss, var_list)` is what the reference's builders return
from `build_subnetwork_train_op` (adanet/examples/simple_dnn.py:103-110) and
what `ComplexityRegularizedEnsembler.build_train_op` returns
(adanet/ensemble/weighted.py:606-617).  Here it yields a `TrainOp` record that
the engine executes with the fused CUDA optimizer kernel (adn_opt_step); the
update rules are TensorFlow 1.x's (see csrc/optim.cu).
"""

from __future__ import annotations

from typing import Optional, Sequence


class TrainOp:
  """A deferred optimizer application: (optimizer spec, loss, var_list)."""

  def __init__(self, kind: str, spec: Optional[tuple] = None, loss=None, var_list: Optional[Sequence] = None):
    self.kind, self.spec, self.loss, self.var_list = kind, spec, loss, list(var_list) if var_list is not None else None
    self.type = "NoOp" if kind == "no_op" else "Minimize"

  def __repr__(self):
    return "TrainOp(%s, %s)" % (self.kind, self.spec)


def no_op(name: Optional[str] = None) -> TrainOp:
  return TrainOp("no_op")


class Optimizer:
  _kind = None

  def spec(self) -> tuple:
    raise NotImplementedError

  def minimize(self, loss, var_list=None, global_step=None) -> TrainOp:
    return TrainOp("minimize", self.spec(), loss, var_list)


class GradientDescentOptimizer(Optimizer):
  """v -= lr * g."""

  def __init__(self, learning_rate):
    self.learning_rate = _constant_lr(learning_rate, "GradientDescentOptimizer")

  def spec(self):
    return ("sgd", self.learning_rate)


class CosineDecay:
  """tf.train.cosine_decay(learning_rate, global_step, decay_steps, alpha) [TF]:
  lr * ((1-alpha) * 0.5 * (1 + cos(pi * min(step, decay_steps)/decay_steps)) + alpha).  `global_step` is the
  builder's `iteration_step` (the only step a builder may schedule on: the Estimator owns the global step,
  customizing_adanet.ipynb SimpleCNNBuilder.build_subnetwork_train_op); the engine evaluates the schedule on the
  device from the optimizer's own step counter, which counts exactly those steps."""

  def __init__(self, learning_rate, global_step, decay_steps, alpha=0.0):
    if float(decay_steps) <= 0:
      raise ValueError("decay_steps must be positive")
    self.learning_rate, self.global_step = float(learning_rate), global_step
    self.decay_steps, self.alpha = float(decay_steps), float(alpha)


def cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0, name=None) -> CosineDecay:
  return CosineDecay(learning_rate, global_step, decay_steps, alpha)


def _constant_lr(learning_rate, who: str) -> float:
  if isinstance(learning_rate, CosineDecay):
    raise NotImplementedError("%s with a cosine_decay learning rate is not implemented (MomentumOptimizer is)" % who)
  return float(learning_rate)


class MomentumOptimizer(Optimizer):
  """acc = momentum*acc + g; v -= lr*acc (use_nesterov=False); lr constant or a `cosine_decay` schedule."""

  def __init__(self, learning_rate, momentum, use_nesterov=False):
    if use_nesterov:
      raise NotImplementedError("Nesterov momentum is not implemented")
    self.schedule = learning_rate if isinstance(learning_rate, CosineDecay) else None
    self.learning_rate = self.schedule.learning_rate if self.schedule else float(learning_rate)
    self.momentum = float(momentum)

  def spec(self):
    if self.schedule is not None:
      return ("momentum_cosine", self.learning_rate, self.momentum, self.schedule.decay_steps, self.schedule.alpha)
    return ("momentum", self.learning_rate, self.momentum)


class RMSPropOptimizer(Optimizer):
  """ms (init 1) = rho*ms + (1-rho)g^2; mom = mu*mom + lr*g/sqrt(ms+eps); v -= mom."""

  def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10, centered=False):
    if centered:
      raise NotImplementedError("centered RMSProp is not implemented")
    self.learning_rate = _constant_lr(learning_rate, "RMSPropOptimizer")
    self.decay, self.momentum, self.epsilon = float(decay), float(momentum), float(epsilon)

  def spec(self):
    return ("rmsprop", self.learning_rate, self.decay, self.momentum, self.epsilon)


class AdamOptimizer(Optimizer):
  """lr_t = lr*sqrt(1-b2^t)/(1-b1^t); var -= lr_t*m/(sqrt(v)+eps)."""

  def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    self.learning_rate = _constant_lr(learning_rate, "AdamOptimizer")
    self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)

  def spec(self):
    return ("adam", self.learning_rate, self.beta1, self.beta2, self.epsilon)


def optimizer_from(obj) -> Optional[tuple]:
  """Accepts an Optimizer, a spec tuple, a zero-arg callable returning either, or None."""
  if obj is None:
    return None
  if callable(obj) and not isinstance(obj, Optimizer):
    obj = obj()
  if isinstance(obj, Optimizer):
    return obj.spec()
  if isinstance(obj, tuple):
    return obj
  raise ValueError("unsupported optimizer %r" % (obj,))
