"""cfl.ops -- the pieces of the reference's cfl/ops.py that touch the hot path (normalisers,
lrelu, small functional helpers).  Image augmentation ops are out of scope."""
from __future__ import annotations

import functools

import torch


def lrelu(x, leak=0.2, name="lrelu"):
    """cfl/ops.py:10-12."""
    return torch.relu(x) - leak * torch.relu(-x)


def relu(x):
    return torch.relu(x)


def tanh(x):
    return torch.tanh(x)


def sigmoid(x):
    return torch.sigmoid(x)


def compose(*functions):
    """cfl/ops.py:21-23."""
    return functools.reduce(lambda f, g: lambda x: f(g(x)), functions, lambda x: x)


def identity(tensor):
    return tensor


def reshapper(shape):
    def fn(tensor):
        return tensor.reshape((-1,) + tuple(shape))
    return fn


def normalize(tensor, scale, shift, clip_value_min=None, clip_value_max=None):
    """cfl/ops.py:198-202."""
    tensor = tensor / scale + shift
    if clip_value_min is not None or clip_value_max is not None:
        tensor = torch.clamp(tensor, clip_value_min, clip_value_max)
    return tensor


class _Normalizer:
    """Callable normaliser that also exposes, when it is a pure scaling, the factor the
    projection kernel can fold in (x*in_scale) instead of materialising the scaled batch."""

    def __init__(self, fn, in_scale=None):
        self.fn, self.in_scale = fn, in_scale

    def __call__(self, tensor):
        return self.fn(tensor)


def normalizer(scale, shift, clip_value_min=None, clip_value_max=None):
    """cfl/ops.py:205-214."""
    pure = shift == 0 and clip_value_min is None and clip_value_max is None
    return _Normalizer(lambda t: normalize(t, scale, shift, clip_value_min, clip_value_max),
                       in_scale=(1.0 / scale) if pure else None)


def normalize_v2(tensor, input_shape, scale=None, mean=None, norm=None, clip_value_min=None,
                 clip_value_max=None):
    """cfl/ops.py:66-124: (x*scale - mean)/norm (per channel when 3 values), clip, flatten."""
    as_tuple = lambda v: v if (v is None or isinstance(v, (list, tuple))) else (v,)
    mean, norm = as_tuple(mean), as_tuple(norm)
    if scale is not None and scale != 1.0:
        tensor = tensor * scale
    if (mean and len(mean) > 1) or (norm and len(norm) > 1):
        chans = list(torch.unbind(tensor, dim=-1))
        if mean:
            m3 = tuple(mean) * 3 if len(mean) == 1 else tuple(mean)
            assert len(m3) == 3
            chans = [c - m if m != 0.0 else c for c, m in zip(chans, m3)]
        if norm:
            n3 = tuple(norm) * 3 if len(norm) == 1 else tuple(norm)
            assert len(n3) == 3
            chans = [c / n if n != 1.0 else c for c, n in zip(chans, n3)]
        tensor = torch.stack(chans, dim=-1)
    else:
        if mean and mean[0] != 0.0:
            tensor = tensor - mean[0]
        if norm and norm[0] != 1.0:
            tensor = tensor / norm[0]
    if clip_value_min is not None and clip_value_max is None:
        tensor = torch.clamp(tensor, min=clip_value_min)
    elif clip_value_min is None and clip_value_max is not None:
        tensor = torch.clamp(tensor, max=clip_value_max)
    elif clip_value_min is not None:
        tensor = torch.clamp(tensor, clip_value_min, clip_value_max)
    if input_shape:
        size = 1
        for dim in input_shape:
            size *= dim
        tensor = tensor.reshape(-1, size)
    return tensor


def normalizer_v2(input_shape, scale=None, mean=None, norm=None, clip_value_min=None, clip_value_max=None):
    """cfl/ops.py:127-143."""
    single = lambda v: v is None or not isinstance(v, (list, tuple)) or len(v) == 1
    first = lambda v: v[0] if isinstance(v, (list, tuple)) else v
    pure = (clip_value_min is None and clip_value_max is None and single(mean) and single(norm)
            and (mean is None or first(mean) == 0))
    in_scale = None
    if pure:
        in_scale = (scale if scale is not None else 1.0) / (first(norm) if norm is not None else 1.0)
    return _Normalizer(lambda t: normalize_v2(t, input_shape, scale, mean, norm, clip_value_min, clip_value_max),
                       in_scale=in_scale)


def unnormalize(tensor, scale, shift):
    """cfl/ops.py:217-218: the inverse of ``normalize`` without its clip: (x - shift) * scale."""
    return (tensor - shift) * scale


def unnormalizer(scale, shift):
    """cfl/ops.py:221-225."""
    return lambda tensor: unnormalize(tensor, scale, shift)


def unnormalize_v2(tensor, input_shape, scale=None, mean=None, norm=None):
    """cfl/ops.py:146-190: the inverse of ``normalize_v2`` without its clip: reshape to ``(-1,) + input_shape``,
    x*norm + mean (per channel when 3 values), then / scale."""
    as_tuple = lambda v: v if (v is None or isinstance(v, (list, tuple))) else (v,)
    mean, norm = as_tuple(mean), as_tuple(norm)
    tensor = tensor.reshape((-1,) + tuple(input_shape))
    if (mean and len(mean) > 1) or (norm and len(norm) > 1):
        chans = list(torch.unbind(tensor, dim=-1))
        if norm:
            n3 = tuple(norm) * 3 if len(norm) == 1 else tuple(norm)
            assert len(n3) == 3
            chans = [c * n if n != 1.0 else c for c, n in zip(chans, n3)]
        if mean:
            m3 = tuple(mean) * 3 if len(mean) == 1 else tuple(mean)
            assert len(m3) == 3
            chans = [c + m if m != 0.0 else c for c, m in zip(chans, m3)]
        tensor = torch.stack(chans, dim=-1)
    else:
        if norm and norm[0] != 1.0:
            tensor = tensor * norm[0]
        if mean and mean[0] != 0.0:
            tensor = tensor + mean[0]
    if scale is not None and scale != 1.0:
        tensor = tensor / scale
    return tensor


def unnormalizer_v2(input_shape, scale=None, mean=None, norm=None):
    """cfl/ops.py:193-198."""
    return lambda tensor: unnormalize_v2(tensor, input_shape, scale=scale, mean=mean, norm=norm)


def dist_normalizer(input_shape, ae_shape, data_scale, data_mean, data_norm, latent_norm, data_type):
    """cfl/ops.py:302-349 -> (data_normalizer, data_unnormalizer, ae_normalizer, ae_unnormalizer,
    latent_normalizer)."""
    clip_values = {"sigmoid": (0.0, 1.0), "tanh": (-1.0, 1.0), "relu": (0.0, None), "linear": (None, None)}
    lo, hi = clip_values[data_type]
    data_normalizer = normalizer_v2(input_shape, scale=data_scale, mean=data_mean, norm=data_norm,
                                    clip_value_min=lo, clip_value_max=hi)
    latent_normalizer = normalizer_v2(None, norm=latent_norm) if latent_norm else None
    data_unnormalizer = unnormalizer_v2(input_shape, scale=data_scale, mean=data_mean, norm=data_norm)
    if ae_shape is not None and tuple(ae_shape) != tuple(input_shape):
        ae_normalizer = normalizer_v2(ae_shape, scale=data_scale, mean=data_mean, norm=data_norm,
                                      clip_value_min=lo, clip_value_max=hi)
        ae_unnormalizer = unnormalizer_v2(ae_shape, scale=data_scale, mean=data_mean, norm=data_norm)
    else:
        ae_normalizer, ae_unnormalizer = data_normalizer, data_unnormalizer
    return data_normalizer, data_unnormalizer, ae_normalizer, ae_unnormalizer, latent_normalizer
