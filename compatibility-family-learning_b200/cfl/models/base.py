"""cfl.models.base -- ModelBase / DistBase with the reference's attribute surface
(cfl/models/base.py), eager: constructing an encoder on a batch X computes its heads with the
sm_100a projection kernel; build_dist(target) returns the [B,1] distance."""
from __future__ import annotations

import torch

from .. import functional as F
from .. import variables as vs
from ..layers import fully_connected_weight_norm


class ModelBase(object):
    """cfl/models/base.py:6-39."""

    def __init__(self, scope):
        self.name = scope.name
        self.scope = scope
        self.summaries = []

    def get_vars(self):
        return list(vs.get_collection(self.name).values())

    def named_vars(self):
        return vs.get_collection(self.name)

    def reg_loss(self):
        """Sum of the regularisation losses under this scope (base.py:16-19):
        l2_regularizer(c) = c * sum(w^2)/2 on every V / weights / biases, never on g."""
        c = getattr(self, "reg_const", 0.0) or 0.0
        if not c:
            return torch.zeros((), device=vs.default_device())
        tot = torch.zeros((), device=vs.default_device())
        for k, v in self.named_vars().items():
            if k.rsplit("/", 1)[-1] in ("V", "weights", "biases"):
                tot = tot + 0.5 * c * (v.detach() ** 2).sum()
        return tot

    def update_ops(self):
        return []

    def add_summary(self, name, op, summary_fn=None, summary_list=None):
        self.add_summary_op((name, op), summary_list=summary_list)

    def add_summary_op(self, op, summary_list=None):
        if summary_list is None:
            summary_list = [self.summaries]
        elif not isinstance(summary_list, list) or len(summary_list) == 0 or not isinstance(summary_list[0], list):
            summary_list = [summary_list]
        for summary in summary_list:
            summary.append(op)


class DistBase(ModelBase):
    """cfl/models/base.py:42-146.  Subclasses set dist_type, num_outputs, num_components,
    regularizer, initializer, gate before calling build_prototypes."""

    in_scale = 1.0

    def build_prototypes(self, flatten_outputs, activation_fn):
        has_bias = self.dist_type.startswith("pcd")            # base.py:45-46, 62-63
        bias_init = "zeros" if has_bias else None
        d, K = self.num_outputs, self.num_components
        with vs.variable_scope("outputs"):
            outputs = fully_connected_weight_norm(flatten_outputs, d, activation_fn=None,
                                                  weights_initializer=self.initializer,
                                                  biases_initializer=bias_init, in_scale=self.in_scale)
            self.outputs = outputs
            self.activations = _apply_act(outputs, activation_fn)
        if self.dist_type in ("pcd", "monomer"):
            with vs.variable_scope("prototype_outputs"):
                proto = fully_connected_weight_norm(flatten_outputs, d * K, activation_fn=None,
                                                    weights_initializer=self.initializer,
                                                    biases_initializer=bias_init, in_scale=self.in_scale)
                proto = _apply_act(proto, activation_fn)
                self.flat_prototype_activations = proto
                self.flat_all_activations = torch.cat([self.activations, proto], dim=-1)
                self.prototype_activations = proto.reshape(-1, K, d)
                if self.gate is not None:
                    idx = self.gate
                    self.one_prototype_activations = self.prototype_activations[idx[:, 0], idx[:, 1]]
                self.all_prototype_activations = list(torch.split(proto, d, dim=1))
        if self.dist_type == "monomer":
            with vs.variable_scope("monomer_outputs"):
                self.monomer_outputs = fully_connected_weight_norm(self.outputs, K, activation_fn=None,
                                                                   weights_initializer=self.initializer,
                                                                   biases_initializer=None)
                self.monomer_activations = torch.softmax(self.monomer_outputs, dim=-1)

    def build_dist(self, target):
        """[B,1] distance to ``target`` (base.py:107-146)."""
        if self.dist_type == "monomer":
            d = F.pair_dist("monomer", self.activations, target.prototype_activations, self.monomer_activations)
        elif self.dist_type == "siamese":
            d = F.pair_dist("siamese", self.activations, target.activations)
        elif self.dist_type.startswith("pcd"):
            d = F.pair_dist("pcd", target.activations, self.prototype_activations)
        else:
            raise ValueError(self.dist_type)
        return d.reshape(-1, 1)


def _apply_act(x, fn):
    if fn is None:
        return x
    if isinstance(fn, str):
        return {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu, "linear": lambda t: t}[fn](x)
    return fn(x)
