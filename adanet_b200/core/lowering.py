"""Lowers a Builder's symbolic Subnetwork to an engine plan spec.

The reference calls `builder.build_subnetwork(...)` inside a TF variable scope
and diffs the variable collections to find the builder's own variables
(adanet/core/ensemble_builder.py:679-805, var_list isolation :224-255,754);
then `builder.build_subnetwork_train_op(subnetwork, loss, var_list, ...)`.
Here the same two calls are made against the symbolic graph of
adanet_b200.graph, and the result -- a chain
input_layer -> [dense+relu]* -> dense -- becomes a SubnetworkPlanSpec.
"""

from __future__ import annotations

import inspect
from typing import Dict, List, Optional, Tuple

import numpy as np

from adanet_b200 import graph
from adanet_b200 import subnetwork as subnetwork_lib
from adanet_b200 import train
from adanet_b200.core import engine as eng


class _NullSummary:
  """Stand-in for adanet.Summary (adanet/core/summary.py): TensorBoard plumbing is out of scope."""

  def scalar(self, name, tensor=None, family=None, **kwargs):
    return None

  image = audio = histogram = scalar

  def current_scope(self):
    import contextlib
    return contextlib.nullcontext()


def call_build_subnetwork(builder, features, labels, logits_dimension, training, iteration_step, summary,
                          previous_ensemble, config=None):
  """Calls build_subnetwork with `labels` / `config` only when the override declares them
  (argument-name sniffing, adanet/core/ensemble_builder.py:737-746)."""
  args = inspect.signature(builder.build_subnetwork).parameters
  kwargs = dict(features=features, logits_dimension=logits_dimension, training=training,
                iteration_step=iteration_step, summary=summary, previous_ensemble=previous_ensemble)
  if "labels" in args:
    kwargs["labels"] = labels
  if "config" in args:
    kwargs["config"] = config
  return builder.build_subnetwork(**kwargs)


def _trace_dense_chain(logits: graph.Tensor) -> Tuple[List[graph.Tensor], graph.Tensor]:
  """Walks logits back to the input layer; returns (dense ops first->last, input tensor)."""
  chain = []
  t = logits
  pending_dropout = None
  while True:
    if t.op == "dropout":
      # tf.layers.dropout on a hidden activation: belongs to the dense layer that produces it
      if pending_dropout is not None or not chain:
        raise NotImplementedError("dropout must follow a hidden dense+relu layer (one dropout per layer)")
      pending_dropout = (t.attrs["rate"], t.attrs["seed"])
      t = t.inputs[0]
    elif t.op == "dense":
      if pending_dropout is not None:
        t.attrs["dropout_after"] = pending_dropout
        pending_dropout = None
      chain.append(t)
      t = t.inputs[0]
    elif t.op == "relu":
      # explicit relu(dense(...)) is folded into the dense's activation
      inner = t.inputs[0]
      if inner.op != "dense" or inner.attrs.get("activation") is not None:
        raise NotImplementedError("relu must directly follow a linear dense layer")
      inner.attrs["activation"] = "relu"
      t = inner
    elif t.op in ("input_layer", "placeholder", "flatten"):
      if pending_dropout is not None:
        raise NotImplementedError("dropout on the input features is not implemented by the B200 engine")
      break
    else:
      raise NotImplementedError(
          "the B200 engine runs dense subnetworks (input_layer -> [dense+relu]* -> dense), optionally behind a "
          "conv3x3+relu -> maxpool2 -> flatten stem; op %r is not supported" % t.op)
  chain.reverse()
  return chain, t


def _trace_conv_stem(flat: graph.Tensor):
  """flatten <- max_pool2d(2,2) <- [relu <-] conv2d(3x3, same) <- NHWC images: the SimpleCNN stem
  (customizing_adanet.ipynb SimpleCNNBuilder.build_subnetwork).  Returns (conv op, image placeholder)."""
  pool = flat.inputs[0]
  if pool.op != "max_pool2d":
    raise NotImplementedError("flatten must follow max_pooling2d in a conv-stem subnetwork (got op %r)" % pool.op)
  conv = pool.inputs[0]
  if conv.op == "relu":
    inner = conv.inputs[0]
    if inner.op != "conv2d" or inner.attrs.get("activation") is not None:
      raise NotImplementedError("relu must directly follow a linear conv2d layer")
    inner.attrs["activation"] = "relu"
    conv = inner
  if conv.op != "conv2d":
    raise NotImplementedError("max_pooling2d must follow conv2d (+relu) in a conv-stem subnetwork (got op %r)" % conv.op)
  if conv.attrs.get("activation") != "relu":
    raise NotImplementedError("the conv stem's activation must be relu")
  images = conv.inputs[0]
  if images.op != "placeholder" or len(images.shape) != 4:
    raise NotImplementedError("the conv stem must read the NHWC image feature directly")
  return conv, images


def lower_subnetwork(builder, sub: subnetwork_lib.Subnetwork, variables: List[graph.Variable], train_op,
                     in_dim: int, logits_dim: int) -> eng.SubnetworkPlanSpec:
  if isinstance(sub.logits, dict):
    raise NotImplementedError("multi-head subnetworks are not implemented by the B200 engine")
  chain, inp = _trace_dense_chain(sub.logits)
  if not chain:
    raise ValueError("subnetwork %s has no dense layer" % builder.name)
  if chain[-1].attrs.get("activation") is not None:
    raise NotImplementedError("the logits layer must be linear")
  for d in chain[:-1]:
    if d.attrs.get("activation") != "relu":
      raise NotImplementedError("hidden layers must use relu")
  if chain[-1].attrs.get("dropout_after") is not None:
    raise NotImplementedError("dropout on the logits is not implemented by the B200 engine")
  stem, image_shape = None, None
  if inp.op == "flatten":
    stem, images = _trace_conv_stem(inp)
    image_shape = tuple(int(v) for v in images.shape[1:])
    if int(np.prod(image_shape)) != in_dim:
      raise ValueError("subnetwork %s consumes images of shape %s, the input_fn provides %d values per example" % (
          builder.name, image_shape, in_dim))
  elif inp.shape[-1] != in_dim or len(inp.shape) != 2:
    raise ValueError("subnetwork %s consumes %s input features, the input_fn provides %d" % (builder.name, inp.shape[-1], in_dim))
  dims = [int(inp.shape[-1])] + [d.shape[-1] for d in chain]
  if dims[-1] != logits_dim:
    raise ValueError("subnetwork %s produces logits of dimension %d, head expects %d" % (builder.name, dims[-1], logits_dim))
  # last_layer must be the tensor feeding the logits layer (or the logits themselves)
  ll = sub.last_layer
  if ll is not chain[-1].inputs[0] and ll is not sub.logits:
    if not (isinstance(ll, graph.Tensor) and ll.op in ("relu", "dropout") and ll.inputs[0] is chain[-1].inputs[0]):
      raise NotImplementedError("last_layer must be the input of the logits layer (or the logits)")
  layers = ([stem] if stem is not None else []) + chain     # a conv stem's HWIO kernel / bias lead the lists
  ws = [np.array(d.attrs["kernel"].value, dtype=np.float32) for d in layers]
  bs = [np.array(d.attrs["bias"].value, dtype=np.float32) if d.attrs["bias"] is not None
        else np.zeros((d.shape[-1],), dtype=np.float32) for d in layers]
  if stem is not None and stem.attrs["bias"] is None:
    raise NotImplementedError("a conv stem without bias is not implemented")
  # train op: TrainOp from optimizer.minimize, or a TrainOpSpec wrapping one
  op = train_op.train_op if isinstance(train_op, subnetwork_lib.TrainOpSpec) else train_op
  if not isinstance(op, train.TrainOp):
    raise ValueError("build_subnetwork_train_op must return optimizer.minimize(...) / TrainOpSpec, got %r" % (op,))
  if op.kind == "no_op":
    opt_spec = ("sgd", 0.0)     # a frozen subnetwork (e.g. estimator_test.py _FrozenLinearBuilder)
  else:
    opt_spec = op.spec
    chain_vars = {id(d.attrs["kernel"]) for d in layers} | {id(d.attrs["bias"]) for d in layers if d.attrs["bias"] is not None}
    if op.var_list is not None and {id(v) for v in op.var_list} != chain_vars:
      raise NotImplementedError("training a strict subset of a subnetwork's variables is not implemented")
  complexity = float(np.asarray(sub.complexity, dtype=np.float32))
  spec = eng.SubnetworkPlanSpec(builder.name, dims, complexity, opt_spec, ws, bs, shared=sub.shared, image_shape=image_shape)
  # which tensor MATRIX mixture weights multiply (weighted.py:449): the logits themselves for sub-estimator builders
  # (autoensemble/common.py:115-118), the penultimate activation for simple_dnn-style builders
  spec.last_layer_is_logits = ll is sub.logits and len(chain) >= 1 and ll is not chain[-1].inputs[0]
  drops = [d.attrs.get("dropout_after") for d in chain[:-1]]
  spec.dropout = drops if any(v is not None for v in drops) else None
  return spec


def build_and_lower(builder, feature_placeholders: Dict[str, graph.Tensor], labels_placeholder, head,
                    iteration_step, previous_ensemble, in_dim: int, config=None):
  """One `_SubnetworkManager.build_subnetwork_spec` (ensemble_builder.py:679-805) against the symbolic graph."""
  summary = _NullSummary()
  with graph.graph_scope() as g:
    sub = call_build_subnetwork(builder, feature_placeholders, labels_placeholder, head.logits_dimension, True,
                                iteration_step, summary, previous_ensemble, config)
    if not isinstance(sub, subnetwork_lib.Subnetwork):
      raise ValueError("build_subnetwork of %s must return an adanet Subnetwork" % builder.name)
    variables = list(g.variables)
    loss = head.create_loss(sub.logits)
    train_op = builder.build_subnetwork_train_op(subnetwork=sub, loss=loss, var_list=variables,
                                                 labels=labels_placeholder, iteration_step=iteration_step,
                                                 summary=summary, previous_ensemble=previous_ensemble)
  spec = lower_subnetwork(builder, sub, variables, train_op, in_dim, head.logits_dimension)
  # bagging: a builder from an AutoEnsembleSubestimator with its own train_input_fn (autoensemble/common.py:151-180)
  spec.own_input = getattr(builder, "bagging_train_input_fn", None) is not None
  return spec, sub
