"""torch.autograd bridges over the C-ABI kernels, so the reference-shaped constructors return
differentiable tensors the way TF graph nodes are differentiable.  The fused train step of the
linear models (cfl/models/_pair_model.py) bypasses autograd and calls the kernels directly."""
from __future__ import annotations

import torch

from . import _native as nat


class _Project(torch.autograd.Function):
    """y = act((in_scale*x @ V) * g/|V_col| + b)  (cfl/layers.py:80-94) or the plain FC."""

    @staticmethod
    def forward(ctx, x, V, g, b, weight_norm, in_scale, act):
        x = x.contiguous()
        y, _, z = nat.project_fwd(x, V, g, b, weight_norm, in_scale, act, want_z=weight_norm)
        ctx.save_for_backward(x, V, g, b, y, z)
        ctx.cfg = (weight_norm, in_scale, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, V, g, b, y, z = ctx.saved_tensors
        weight_norm, in_scale, act = ctx.cfg
        dy = dy.contiguous()
        dV, dg, db = nat.project_bwd(x, V, g, b, weight_norm, in_scale, act, y, z, dy,
                                     want_dg=g is not None, want_dbias=b is not None)
        dx = None
        if ctx.needs_input_grad[0]:
            # gradient w.r.t. the layer input is only needed when a trunk sits in front of the
            # heads (ConvPCD / hidden FC layers): small plumbing matmul, not on the named hot path
            dpre = dy
            if act not in (None, "linear"):
                dpre = dy * _act_grad(y, act)
            Vs = V
            if weight_norm:
                s = (g if g is not None else 1.0) / V.pow(2).sum(0).sqrt()
                Vs = V * s
            dx = (dpre @ Vs.t()) * in_scale
        return dx, dV, (dg if g is not None else None), (db if b is not None else None), None, None, None


def _act_grad(y, act):
    if act == "tanh":
        return 1 - y * y
    if act == "sigmoid":
        return y * (1 - y)
    if act == "relu":
        return (y > 0).to(y.dtype)
    if act == "lrelu":
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.2))
    return torch.ones_like(y)


def project(x, V, g=None, b=None, weight_norm=True, in_scale=1.0, act=None):
    return _Project.apply(x, V, g, b, weight_norm, in_scale, act)


class _PairDist(torch.autograd.Function):
    """DistBase.build_dist (cfl/models/base.py:107-146) on paired rows -> dist [B]."""

    @staticmethod
    def forward(ctx, mode, a, P, w):
        a, P = a.contiguous(), P.contiguous()
        dist, _, _, _ = nat.pair_loss_fwd(mode, a, P, w=w)
        ctx.save_for_backward(a, P, w)
        ctx.mode = mode
        return dist

    @staticmethod
    def backward(ctx, ddist):
        a, P, w = ctx.saved_tensors
        da, dP, dw, _ = nat.pair_loss_bwd(ctx.mode, a, P, w=w, ddist=ddist.contiguous())
        return None, da, dP.reshape(P.shape), dw


def pair_dist(mode, a, P, w=None):
    if P.dim() == 2:
        P = P[:, None, :]
    return _PairDist.apply(mode, a, P, w)
