"""cfl.layers -- same names and argument meaning as the reference (cfl/layers.py), eager PyTorch +
the sm_100a projection kernel instead of TF graph ops."""
from __future__ import annotations

import torch

from . import functional as F
from . import variables as vs

_ACT_NAMES = {None: None, "linear": None, "tanh": "tanh", "sigmoid": "sigmoid", "relu": "relu", "lrelu": "lrelu"}


def _act_name(fn):
    """Maps an activation given as a name or as one of the cfl.ops callables to the kernel's enum."""
    if fn is None or isinstance(fn, str):
        return _ACT_NAMES[fn]
    name = getattr(fn, "__name__", "")
    if name in ("relu", "tanh", "sigmoid", "lrelu"):
        return name
    raise ValueError(f"unsupported activation_fn {fn!r} (use None/'tanh'/'sigmoid'/'relu'/cfl.ops.lrelu)")


def fully_connected_weight_norm(inputs, num_outputs, activation_fn="relu", weights_initializer=None,
                                weights_regularizer=None, biases_initializer="zeros",
                                biases_regularizer=None, scale=True, reuse=None,
                                variables_collections=None, outputs_collections=None, trainable=True,
                                scope=None, in_scale=1.0):
    """cfl/layers.py:28-97: variables ``g`` (ones), ``V`` (Xavier), ``biases`` (zeros, optional)
    under ``<scope or 'fully_connected'>``; y = act((x @ V) * g/sqrt(sum(V^2,0)) + biases).
    ``biases_initializer=None`` drops the bias (cfl/models/base.py:45-46).  ``in_scale`` folds
    the input normaliser x/scale into the kernel (cfl/ops.py:198)."""
    if not isinstance(num_outputs, int):
        raise ValueError("num_outputs should be int or long, got %s." % (num_outputs,))
    if inputs.dim() != 2:
        inputs = inputs.reshape(inputs.shape[0], -1)
    with vs.variable_scope(scope, "fully_connected", reuse=reuse):
        g = vs.get_variable("g", [num_outputs], vs.ones_initializer()) if scale else None
        V = vs.get_variable("V", [inputs.shape[1], num_outputs], weights_initializer or vs.xavier_initializer())
        b = None
        if biases_initializer is not None:
            init = vs.zeros_initializer() if biases_initializer == "zeros" else biases_initializer
            b = vs.get_variable("biases", [num_outputs], init)
    return F.project(inputs, V, g, b, True, in_scale, _act_name(activation_fn))


def fully_connected(inputs, num_outputs, activation_fn=None, weights_initializer=None,
                    weights_regularizer=None, biases_initializer="zeros", biases_regularizer=None,
                    reuse=None, scope=None, in_scale=1.0):
    """tf.contrib.layers.fully_connected as used by FCEncoder (cfl/models/dist.py:45-65):
    variables ``weights`` and ``biases`` under ``fully_connected``."""
    if inputs.dim() != 2:
        inputs = inputs.reshape(inputs.shape[0], -1)
    with vs.variable_scope(scope, "fully_connected", reuse=reuse):
        W = vs.get_variable("weights", [inputs.shape[1], num_outputs], weights_initializer or vs.xavier_initializer())
        b = None
        if biases_initializer is not None:
            init = vs.zeros_initializer() if biases_initializer == "zeros" else biases_initializer
            b = vs.get_variable("biases", [num_outputs], init)
    return F.project(inputs, W, None, b, False, in_scale, _act_name(activation_fn))


def conv2d_weight_norm(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn="relu",
                       weights_initializer=None, weights_regularizer=None, biases_initializer="zeros",
                       scale=True, reuse=None, scope=None):
    """cfl/layers.py:100-187 (NHWC): W = g * l2_normalize(V, [0,1,2]) (eps 1e-12), conv, bias,
    activation.  Plumbing for the ConvPCD trunk (BASELINE config 1): torch conv2d, asymmetric SAME
    padding as TF."""
    import torch.nn.functional as TF
    from . import ops
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    cin = inputs.shape[-1]
    with vs.variable_scope(scope, "Conv", reuse=reuse):
        g = vs.get_variable("g", [num_outputs], vs.ones_initializer()) if scale else None
        V = vs.get_variable("V", [kh, kw, cin, num_outputs], weights_initializer or vs.xavier_initializer())
        b = vs.get_variable("biases", [num_outputs], vs.zeros_initializer()) if biases_initializer is not None else None
    norm = torch.sqrt(torch.clamp((V * V).sum(dim=(0, 1, 2), keepdim=True), min=1e-12))
    W = V / norm
    if g is not None:
        W = W * g.view(1, 1, 1, -1)
    x = inputs.permute(0, 3, 1, 2)                           # NHWC -> NCHW
    if padding == "SAME":
        H, Wd = x.shape[2], x.shape[3]
        ph = max((-(-H // sh) - 1) * sh + kh - H, 0)
        pw = max((-(-Wd // sw) - 1) * sw + kw - Wd, 0)
        x = TF.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    # fp32 convolution: cuDNN's TF32 default (10-bit significand) would miss the 1e-4 contract of the path by 10x
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        y = TF.conv2d(x, W.permute(3, 2, 0, 1), bias=b, stride=(sh, sw))
    y = y.permute(0, 2, 3, 1)
    act = _act_name(activation_fn)
    if act == "lrelu":
        y = ops.lrelu(y)
    elif act is not None:
        y = getattr(torch, act)(y)
    return y
