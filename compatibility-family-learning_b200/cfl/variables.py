"""Minimal stand-in for tf.variable_scope / tf.get_variable (TF-1.x), enough for the hot-path
constructors to keep the reference's variable names, e.g.
``CFL/DistEncoder/outputs/fully_connected/V`` or ``Dist/Encoder/pcd_outputs/fully_connected/weights``.
Variables are plain CUDA tensors kept in a process-wide store; ``reuse=True`` fetches them.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List, Optional

import torch

_STORE: Dict[str, torch.Tensor] = {}
_SCOPE: List[str] = []
_REUSE: List[bool] = []
_DEVICE = [None]


def set_default_device(device):
    _DEVICE[0] = torch.device(device) if device is not None else None


def default_device():
    if _DEVICE[0] is not None:
        return _DEVICE[0]
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def reset_default_graph():
    """tf.reset_default_graph(): forget every variable."""
    _STORE.clear()
    _SCOPE.clear()
    _REUSE.clear()


class Scope:
    def __init__(self, name: str):
        self.name = name

    def __repr__(self):
        return f"Scope({self.name!r})"


@contextlib.contextmanager
def variable_scope(name: Optional[str], default_name: Optional[str] = None, reuse: Optional[bool] = None):
    part = name if name is not None else default_name
    if isinstance(part, Scope):          # re-entering a captured scope: absolute name
        saved = list(_SCOPE)
        _SCOPE[:] = part.name.split("/") if part.name else []
    else:
        saved = None
        _SCOPE.append(str(part))
    _REUSE.append(bool(reuse) or (bool(_REUSE) and _REUSE[-1]))
    try:
        yield Scope("/".join(_SCOPE))
    finally:
        _REUSE.pop()
        if saved is not None:
            _SCOPE[:] = saved
        else:
            _SCOPE.pop()


def current_scope() -> str:
    return "/".join(_SCOPE)


def get_variable(name: str, shape, initializer: Callable, dtype=torch.float32) -> torch.Tensor:
    full = "/".join(_SCOPE + [name])
    reuse = bool(_REUSE) and _REUSE[-1]
    if full in _STORE:
        if not reuse:
            raise ValueError(f"Variable {full} already exists, disallowed. Did you mean to set reuse=True?")
        return _STORE[full]
    if reuse:
        raise ValueError(f"Variable {full} does not exist, or was not created with get_variable()")
    t = initializer(tuple(shape)).to(device=default_device(), dtype=dtype).contiguous()
    t.requires_grad_(True)                 # leaf: the autograd path (conv trunk) trains it too
    _STORE[full] = t
    return t


def get_collection(scope: str) -> Dict[str, torch.Tensor]:
    """tf.get_collection(TRAINABLE_VARIABLES, scope=...): variables whose name starts with scope."""
    pre = scope.rstrip("/") + "/"
    return {k: v for k, v in _STORE.items() if k.startswith(pre)}


def all_variables() -> Dict[str, torch.Tensor]:
    return dict(_STORE)


# ---- initializers (tf.contrib.layers.xavier_initializer / zeros / ones / constant) --------------
_GEN = [None]


def set_seed(seed: int):
    _GEN[0] = torch.Generator().manual_seed(int(seed))


def xavier_initializer():
    def init(shape):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (
            int(torch.tensor(shape[:-1]).prod()), shape[-1] * int(torch.tensor(shape[:-2]).prod()) if len(shape) > 2 else shape[-1])
        if len(shape) == 4:                      # conv kernels [kh, kw, cin, cout]
            rf = shape[0] * shape[1]
            fan_in, fan_out = rf * shape[2], rf * shape[3]
        lim = (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(shape, generator=_GEN[0]) * 2 - 1) * lim
    return init


def zeros_initializer():
    return lambda shape: torch.zeros(shape)


def ones_initializer():
    return lambda shape: torch.ones(shape)


def constant_initializer(value):
    return lambda shape: torch.full(shape, float(value))
