"""The simple DNN search space (mirror of adanet/examples/simple_dnn.py).

`_SimpleDNNBuilder` (:26-131): input_layer -> num_layers x [dense(layer_size)+relu
(+dropout)] -> dense(logits); complexity sqrt(num_layers); shared {"num_layers"};
named "linear" / "{n}_layer_dnn".  `Generator` (:134-213) proposes two builders
per iteration: as deep as the most recent subnetwork, and one layer deeper.
"""

from __future__ import annotations

import functools
import math

import adanet_b200 as adanet
from adanet_b200 import graph
from adanet_b200 import train

_NUM_LAYERS_KEY = "num_layers"


class _SimpleDNNBuilder(adanet.subnetwork.Builder):
  """Builds a DNN subnetwork for AdaNet."""

  def __init__(self, feature_columns, optimizer, layer_size, num_layers, learn_mixture_weights, dropout, seed):
    self._feature_columns = feature_columns
    self._optimizer = optimizer
    self._layer_size = layer_size
    self._num_layers = num_layers
    self._learn_mixture_weights = learn_mixture_weights
    self._dropout = dropout
    self._seed = seed

  def build_subnetwork(self, features, logits_dimension, training, iteration_step, summary, previous_ensemble=None):
    input_layer = graph.input_layer(features=features, feature_columns=self._feature_columns)
    last_layer = input_layer
    for _ in range(self._num_layers):
      last_layer = graph.dense(last_layer, units=self._layer_size, activation=graph.relu,
                               kernel_initializer=graph.glorot_uniform_initializer(seed=self._seed))
      last_layer = graph.dropout(last_layer, rate=self._dropout, seed=self._seed, training=training)
    logits = graph.dense(last_layer, units=logits_dimension,
                         kernel_initializer=graph.glorot_uniform_initializer(seed=self._seed))
    # Rademacher complexity approximated by sqrt(depth) (simple_dnn.py:88-90), in fp32 like tf.sqrt
    complexity = float(math.sqrt(self._num_layers))
    summary.scalar("complexity", complexity)
    summary.scalar("num_layers", self._num_layers)
    shared = {_NUM_LAYERS_KEY: self._num_layers}
    return adanet.Subnetwork(last_layer=last_layer, logits=logits, complexity=complexity, shared=shared)

  def build_subnetwork_train_op(self, subnetwork, loss, var_list, labels, iteration_step, summary, previous_ensemble):
    return self._optimizer.minimize(loss=loss, var_list=var_list)

  def build_mixture_weights_train_op(self, loss, var_list, logits, labels, iteration_step, summary):
    """Deprecated path (simple_dnn.py:112-122); the Ensembler's build_train_op is used instead."""
    if not self._learn_mixture_weights:
      return train.no_op("mixture_weights_train_op")
    return self._optimizer.minimize(loss=loss, var_list=var_list)

  @property
  def name(self):
    if self._num_layers == 0:
      return "linear"     # a DNN with no hidden layers is a linear model
    return "{}_layer_dnn".format(self._num_layers)


class Generator(adanet.subnetwork.Generator):
  """Generates two DNN subnetworks at each iteration (simple_dnn.py:134-213)."""

  def __init__(self, feature_columns, optimizer, layer_size=32, initial_num_layers=0, learn_mixture_weights=False,
               dropout=0., seed=None):
    if not feature_columns:
      raise ValueError("feature_columns must not be empty")
    if layer_size < 1:
      raise ValueError("layer_size must be >= 1")
    if initial_num_layers < 0:
      raise ValueError("initial_num_layers must be >= 0")
    self._initial_num_layers = initial_num_layers
    self._dnn_builder_fn = functools.partial(_SimpleDNNBuilder, feature_columns=feature_columns, optimizer=optimizer,
                                             layer_size=layer_size, learn_mixture_weights=learn_mixture_weights,
                                             dropout=dropout, seed=seed)

  def generate_candidates(self, previous_ensemble, iteration_number, previous_ensemble_reports, all_reports):
    num_layers = self._initial_num_layers
    if previous_ensemble:
      num_layers = previous_ensemble.weighted_subnetworks[-1].subnetwork.shared[_NUM_LAYERS_KEY]
    return [self._dnn_builder_fn(num_layers=num_layers), self._dnn_builder_fn(num_layers=num_layers + 1)]
