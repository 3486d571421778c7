"""cfl.models.blocks -- Thresholder, FCPCD, ConvPCD (cfl/models/blocks.py:13-22, 477-590).  The
GAN blocks of the reference are out of scope."""
from __future__ import annotations

import torch

from .. import variables as vs
from ..layers import conv2d_weight_norm, fully_connected_weight_norm
from ..ops import lrelu
from .base import DistBase, ModelBase


class Thresholder(ModelBase):
    """score = -X + max(threshold, 1e-6); threshold initialised AT 1e-6 (blocks.py:13-22)."""

    def __init__(self, X, name="Thresholder", reuse=False):
        with vs.variable_scope(name, reuse=reuse) as scope:
            super().__init__(scope)
            with vs.variable_scope("threshold"):
                self.raw_threshold = vs.get_variable("threshold", [], vs.constant_initializer(1e-6))
                # clamp(min=) routes the tie gradient to the variable like tf.maximum does
                self.threshold = torch.clamp(self.raw_threshold, min=1e-6)
            self.outputs = -1.0 * X + self.threshold


class FCPCD(DistBase):
    """Encode vectors to a latent space and produce `num_components` prototype vectors
    (blocks.py:477-527)."""

    def __init__(self, X, num_outputs, num_components, input_shape, batch_size, gate=None,
                 dist_type="pcd", layer_sizes=None, initializer=None, regularizer=None,
                 layer_activation_fn=lrelu, activation_fn=None, name="encoder", reuse=False,
                 in_scale=1.0):
        self.input_shape = input_shape
        self.num_components = num_components
        self.num_outputs = num_outputs
        self.batch_size = batch_size
        self.regularizer = regularizer
        self.reg_const = regularizer or 0.0
        self.initializer = initializer
        self.dist_type = dist_type
        self.gate = gate
        self.in_scale = in_scale
        layer_sizes = layer_sizes or []
        with vs.variable_scope(name, reuse=reuse) as scope:
            super().__init__(scope)
            outputs = X
            for i, layer_size in enumerate(layer_sizes):
                with vs.variable_scope("fc_{}".format(i)):
                    outputs = fully_connected_weight_norm(outputs, layer_size, activation_fn=layer_activation_fn,
                                                          weights_initializer=initializer,
                                                          in_scale=in_scale if i == 0 else 1.0)
            if layer_sizes:
                self.in_scale = 1.0
            self.build_prototypes(outputs, activation_fn)


class ConvPCD(DistBase):
    """Conv trunk (5x5 stride-2 weight-norm convs + lrelu) in front of the heads
    (blocks.py:530-590).  The trunk is torch plumbing; the heads run on the projection kernel."""

    def __init__(self, X, input_shape, num_components, num_outputs, batch_size, gate=None, dim=64,
                 max_dim=512, min_dim=4, dist_type="pcd", initializer=None, regularizer=None,
                 layer_activation_fn=lrelu, activation_fn=None, name="encoder", reuse=False):
        self.input_shape = input_shape
        self.num_components = num_components
        self.num_outputs = num_outputs
        self.batch_size = batch_size
        self.regularizer = regularizer
        self.reg_const = regularizer or 0.0
        self.initializer = initializer
        self.dist_type = dist_type
        self.gate = gate
        start_dim = min(input_shape[0], input_shape[1])
        nb_upconv = 0
        while start_dim % 2 == 0 and start_dim > min_dim:
            start_dim //= 2
            nb_upconv += 1
        with vs.variable_scope(name, reuse=reuse) as scope:
            super().__init__(scope)
            outputs = X.reshape((-1,) + tuple(input_shape))
            for i in range(nb_upconv):
                with vs.variable_scope("conv{}".format(i + 1)):
                    outputs = conv2d_weight_norm(outputs, dim, kernel_size=(5, 5), stride=(2, 2), padding="SAME",
                                                 activation_fn=lrelu, weights_initializer=initializer)
                dim = min(dim * 2, max_dim)
            flatten_outputs = outputs.reshape(outputs.shape[0], -1)
            self.build_prototypes(flatten_outputs, activation_fn)
