"""cfl.models.dist -- FCEncoder and the ``Dist`` model ("linear_dist", the Monomer-style model of
train_dist / predict_dist; cfl/models/dist.py).  Same constructor arguments; eager execution on
the sm_100a kernels instead of a TF graph."""
from __future__ import annotations

from types import SimpleNamespace

from .. import functional as F
from .. import variables as vs
from ..layers import fully_connected
from ..ops import lrelu
from ..utils import reduce_product
from ._pair_model import PairModel
from .base import ModelBase
from .blocks import Thresholder  # noqa: F401  (re-exported like the reference)


class FCEncoder(ModelBase):
    """Plain FC heads: latent_outputs [B,d], pcd_outputs [B,K,d] (dist.py:12-68)."""

    def __init__(self, X, input_shape, num_components, num_outputs, batch_size, initializer=None,
                 regularizer=None, layer_activation_fn=lrelu, activation_fn=None, name="Encoder",
                 reuse=False, in_scale=1.0):
        self.input_shape = input_shape
        self.num_components = num_components
        self.num_outputs = num_outputs
        self.batch_size = batch_size
        self.regularizer = regularizer
        self.reg_const = regularizer or 0.0
        self.initializer = initializer
        with vs.variable_scope(name, reuse=reuse) as scope:
            super().__init__(scope)
            with vs.variable_scope("latent_outputs"):
                self.latent_outputs = fully_connected(X, num_outputs, activation_fn=activation_fn,
                                                      weights_initializer=initializer, in_scale=in_scale)
            with vs.variable_scope("pcd_outputs"):
                pcd = fully_connected(X, num_outputs * num_components, activation_fn=activation_fn,
                                      weights_initializer=initializer, in_scale=in_scale)
                self.pcd_outputs = pcd.reshape(-1, num_components, num_outputs)

    def build_dist(self, target):
        """dist.py:70-89 (same arithmetic as DistBase.build_dist, pcd branch)."""
        return F.pair_dist("pcd", target.latent_outputs, self.pcd_outputs).reshape(-1, 1)


class Dist(PairModel):
    """dist.py:92-327.  ``train_step(src_pos, dst_pos, src_neg, dst_neg)`` replaces
    ``sess.run(model.s_optim, ...)``; ``predict(src, dst)`` replaces
    ``sess.run(model.val_s_pos_predicts.outputs, feed)``."""

    def __init__(self, input_shape, latent_size, num_components, batch_size, lr, beta1, beta2,
                 batches=None, val_batches=None, normalize_value=None, data_normalizer=None,
                 data_unnormalizer=None, reg_const=0.0, name="Dist", run_tag=None, reuse=False):
        self.is_double = False
        self.input_shape = tuple(input_shape)
        self.batch_size = batch_size
        self.normalize_value = normalize_value
        self.run_tag = run_tag
        self.data_normalizer = data_normalizer
        self.data_unnormalizer = data_unnormalizer
        in_scale = getattr(data_normalizer, "in_scale", None)
        if data_normalizer is not None and in_scale is None:
            raise ValueError("Dist: only pure-scaling normalisers (cfl.ops.normalizer(scale, 0)) can be folded "
                             "into the projection kernel")
        with vs.variable_scope(name, reuse=reuse) as scope:
            ModelBase.__init__(self, scope)
            self._init_pair_model(
                input_size=reduce_product(input_shape), latent_size=latent_size, num_components=num_components,
                dist_type="pcd", act_type=None, weight_norm=False, pos_weight=None, use_threshold=True,
                caffe_margin=None, lambda_m=None, reg_const=reg_const, directed=False, lr=lr, beta1=beta1,
                beta2=beta2, in_scale=in_scale if in_scale is not None else 1.0,
                head_names=("latent_outputs", "pcd_outputs"), encoder_names=("Encoder", "Encoder"))
        self.thres_loss = None

    def get_name(self):
        """dist.py:192-199 (checkpoint / predict directory name)."""
        name = "linear_dist"
        name += "_ls_{}_nc_{}_reg_{}_norm_{}".format(self.latent_size, self.num_components, self.reg_const,
                                                     self.normalize_value)
        if self.run_tag:
            name += "_run_" + self.run_tag
        return name

    def train_step(self, *a, **kw):
        out = super().train_step(*a, **kw)
        self.thres_loss = out["s_thres_loss"]
        return out


def construct_model(input_shape, latent_size, num_components, batch_size, lr, beta1, beta2,
                    normalize_value=None, reg_const=0.0, run_tag=None, **unused):
    """dist.py:330-459 minus the TF queues: returns (model, aux)."""
    from ..ops import normalizer
    data_normalizer = normalizer(normalize_value, 0.0) if normalize_value else None
    model = Dist(input_shape=input_shape, latent_size=latent_size, num_components=num_components,
                 batch_size=batch_size, lr=lr, beta1=beta1, beta2=beta2, normalize_value=normalize_value,
                 data_normalizer=data_normalizer, reg_const=reg_const, run_tag=run_tag)
    return model, SimpleNamespace(queue=None, enqueue_op=None)
