"""An eager stand-in for the slice of the TensorFlow 1.x API that the reference's distance-model
code calls, so that the REFERENCE'S OWN PYTHON (cfl/layers.py, cfl/models/{base,blocks,cfl,dist}.py,
cfl/ops.py, cfl/input_data.py) can execute in this container, where TensorFlow is not installed, and
emit golden vectors (make_reference_golden.py).

What this pins: every line of graph-building code in the reference -- which tensor is source and
which target, axes of every reduction, scope / variable names, reuse flags, which variables are
regularised and optimised, the loss wiring.  What it cannot pin: the semantics of the TF primitives
themselves, which are restated here from TF 1.x's documented behaviour on float64 torch tensors
(``tf.maximum`` routes the gradient to its first argument where x >= y; ``l2_regularizer(c)`` is
c * sum(w^2) / 2; ``sigmoid_cross_entropy_with_logits`` is max(x,0) - x z + log1p(exp(-|x|));
``variable_scope(reuse=True)`` is inherited by inner scopes; ``get_collection(key, scope)`` is a
prefix match).  Test infrastructure only.
"""
import contextlib
import sys
import types

import numpy as np
import torch

DT = torch.float64


class DType(object):
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    @property
    def base_dtype(self):
        return self

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", DT)       # computed in float64: the fixtures are the exact-math answers
float64 = DType("float64", DT)
int32 = DType("int32", torch.int64)
int64 = DType("int64", torch.int64)
bool_ = DType("bool", torch.bool)


def _raw(x):
    if isinstance(x, T):
        return x.v
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, T) for e in x):
        return torch.stack([_raw(e) for e in x])
    a = np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a, dtype=DT)
    return torch.as_tensor(a)


class T(object):
    """Eager tensor with the handful of Tensor methods the reference touches."""
    __array_priority__ = 1000

    def __init__(self, v, name=None):
        self.v = _raw(v)
        self.name = name

    @property
    def dtype(self):
        if self.v.dtype == torch.bool:
            return bool_
        return float32 if self.v.dtype.is_floating_point else int32

    @property
    def shape(self):
        return tuple(self.v.shape)

    def get_shape(self):
        return list(self.v.shape)

    def numpy(self):
        return self.v.detach().numpy()

    def __add__(self, o): return T(self.v + _raw(o))
    def __radd__(self, o): return T(_raw(o) + self.v)
    def __sub__(self, o): return T(self.v - _raw(o))
    def __rsub__(self, o): return T(_raw(o) - self.v)
    def __mul__(self, o): return T(self.v * _raw(o))
    def __rmul__(self, o): return T(_raw(o) * self.v)
    def __truediv__(self, o): return T(self.v / _raw(o))
    def __rtruediv__(self, o): return T(_raw(o) / self.v)
    def __neg__(self): return T(-self.v)
    def __getitem__(self, i): return T(self.v[i])


class Variable(T):
    pass


# ---- graph state --------------------------------------------------------------------------------
class GraphKeys(object):
    TRAINABLE_VARIABLES = "trainable_variables"
    GLOBAL_VARIABLES = "variables"
    REGULARIZATION_LOSSES = "regularization_losses"
    UPDATE_OPS = "update_ops"
    SUMMARIES = "summaries"


class VarScope(object):
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse
        self.original_name_scope = name + "/"


class _State(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.scopes = [VarScope("", False)]
        self.variables = {}
        self.collections = {}
        self.presets = {}
        self.created = []
        self.minimize_ops = []


STATE = _State()


def reset_default_graph():
    STATE.reset()


def preset_variables(values):
    """Values by TF name (without ':0') that get_variable uses instead of the initializer."""
    STATE.presets = {k: np.asarray(v, dtype=np.float64) for k, v in values.items()}


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None):
    cur = STATE.scopes[-1]
    if isinstance(name_or_scope, VarScope):
        name = name_or_scope.name
    else:
        leaf = name_or_scope if name_or_scope is not None else default_name
        name = cur.name + "/" + leaf if cur.name else leaf
    sc = VarScope(name, True if reuse else cur.reuse)
    STATE.scopes.append(sc)
    try:
        yield sc
    finally:
        STATE.scopes.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name + "/"


def add_to_collection(key, value):
    STATE.collections.setdefault(key, []).append(value)


def get_collection(key, scope=None):
    items = STATE.collections.get(key, [])
    if scope is None:
        return list(items)
    return [i for i in items if getattr(i, "name", None) and i.name.startswith(scope)]


def get_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, **unused):
    sc = STATE.scopes[-1]
    full = sc.name + "/" + name if sc.name else name
    if full in STATE.variables:
        if not sc.reuse:
            raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True?" % full)
        return STATE.variables[full]
    if sc.reuse:
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()." % full)
    shape = [int(s) for s in shape]
    if full in STATE.presets:
        val = torch.as_tensor(STATE.presets[full], dtype=DT).reshape(shape).clone()
    else:
        val = _raw(initializer(shape)).to(DT).reshape(shape).clone()
    var = Variable(val.requires_grad_(True), name=full + ":0")
    STATE.variables[full] = var
    STATE.created.append(full)
    add_to_collection(GraphKeys.GLOBAL_VARIABLES, var)
    if trainable:
        add_to_collection(GraphKeys.TRAINABLE_VARIABLES, var)
    if regularizer is not None:
        loss = regularizer(var)
        if loss is not None:
            loss.name = full + "/Regularizer/l2_regularizer:0"
            add_to_collection(GraphKeys.REGULARIZATION_LOSSES, loss)
    return var


# ---- initializers / regularizers -------------------------------------------------------------------
def zeros_initializer(*a, **k):
    return lambda shape, **kw: torch.zeros(list(shape), dtype=DT)


def ones_initializer(*a, **k):
    return lambda shape, **kw: torch.ones(list(shape), dtype=DT)


def constant_initializer(value=0.0, *a, **k):
    return lambda shape, **kw: torch.full(list(shape), float(value), dtype=DT)


def xavier_initializer(uniform=True, seed=None, dtype=None):
    def init(shape, **kw):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (int(np.prod(shape[:-1])), shape[-1])
        lim = (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(list(shape), dtype=DT) * 2 - 1) * lim
    return init


def l2_regularizer(scale, scope=None):
    if scale == 0.0:
        return lambda _: None
    return lambda w: T(scale * (w.v ** 2).sum() / 2)


# ---- ops -----------------------------------------------------------------------------------------
def _axis(axis):
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return axis


def reduce_sum(x, axis=None, keep_dims=False, keepdims=False, name=None):
    v = _raw(x)
    return T(v.sum() if axis is None else v.sum(dim=_axis(axis), keepdim=keep_dims or keepdims))


def reduce_mean(x, axis=None, keep_dims=False, keepdims=False, name=None):
    v = _raw(x)
    return T(v.mean() if axis is None else v.mean(dim=_axis(axis), keepdim=keep_dims or keepdims))


def reshape(x, shape, name=None):
    return T(_raw(x).reshape([int(s) for s in shape]))


def square(x, name=None): return T(_raw(x) ** 2)
def sqrt(x, name=None): return T(torch.sqrt(_raw(x)))
def subtract(a, b, name=None): return T(_raw(a) - _raw(b))
def multiply(a, b, name=None): return T(_raw(a) * _raw(b))
def matmul(a, b, name=None): return T(_raw(a) @ _raw(b))
def concat(xs, axis, name=None): return T(torch.cat([_raw(x) for x in xs], dim=axis))
def stack(xs, axis=0, name=None): return T(torch.stack([_raw(x) for x in xs], dim=axis))
def unstack(x, axis=0, name=None): return [T(t) for t in torch.unbind(_raw(x), dim=axis)]
def split(x, n, axis=0, name=None): return [T(t) for t in torch.chunk(_raw(x), n, dim=axis)]
def expand_dims(x, axis, name=None): return T(_raw(x).unsqueeze(axis))
def transpose(x, perm=None, name=None): return T(_raw(x).permute(*perm) if perm else _raw(x).t())
def tile(x, multiples, name=None): return T(_raw(x).repeat(*multiples))
def ones_like(x, name=None): return T(torch.ones_like(_raw(x)))
def zeros_like(x, name=None): return T(torch.zeros_like(_raw(x)))
def constant(v, dtype=None, name=None): return T(v)
def greater(a, b, name=None): return T(_raw(a) > _raw(b))
def less_equal(a, b, name=None): return T(_raw(a) <= _raw(b))
def cast(x, dtype, name=None): return T(_raw(x).to(dtype.torch))
def group(*ops, **k): return list(ops)
def set_random_seed(seed): torch.manual_seed(seed)
def range_(n, *a, **k): return T(torch.arange(int(n)))


def add_n(xs, name=None):
    out = _raw(xs[0])
    for x in xs[1:]:
        out = out + _raw(x)
    return T(out)


def maximum(x, y, name=None):
    x, y = torch.broadcast_tensors(_raw(x).to(DT), _raw(y).to(DT))
    return T(torch.where(x >= y, x, y))            # TF _MaximumGrad: d/dx where x >= y


def minimum(x, y, name=None):
    x, y = torch.broadcast_tensors(_raw(x).to(DT), _raw(y).to(DT))
    return T(torch.where(x <= y, x, y))


def clip_by_value(t, lo, hi, name=None):
    return minimum(maximum(t, lo), hi)


def gather_nd(params, indices, name=None):
    p, i = _raw(params), _raw(indices).long()
    return T(p[tuple(i[:, k] for k in range(i.shape[1]))])


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    if dtype in (int32, int64):
        return T(torch.randint(int(minval), int(maxval), list(shape)))
    hi = 1.0 if maxval is None else maxval
    return T(torch.rand(list(shape), dtype=DT) * (hi - minval) + minval)


def placeholder(dtype, shape=None, name=None):
    return T(torch.zeros(0, dtype=DT), name=name)          # never fed here: generators pass every batch


def placeholder_with_default(default, shape, name=None):
    return default if isinstance(default, T) else T(default, name=name)


# ---- tf.nn ----------------------------------------------------------------------------------------
def relu(x, name=None): return T(torch.relu(_raw(x)))
def tanh(x, name=None): return T(torch.tanh(_raw(x)))
def sigmoid(x, name=None): return T(torch.sigmoid(_raw(x)))
def softmax(x, dim=-1, name=None): return T(torch.softmax(_raw(x), dim=dim))
def bias_add(x, b, name=None): return T(_raw(x) + _raw(b))
def l2_loss(x, name=None): return T((_raw(x) ** 2).sum() / 2)


def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    x, z = _raw(logits), _raw(labels)
    return T(torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x))))


def fully_connected(inputs, num_outputs, activation_fn=relu, weights_initializer=None, weights_regularizer=None,
                    biases_initializer=zeros_initializer(), biases_regularizer=None, reuse=None, scope=None,
                    trainable=True, **unused):
    """tf.contrib.layers.fully_connected: variables ``weights`` / ``biases`` under scope 'fully_connected'."""
    with variable_scope(scope, "fully_connected", reuse=reuse):
        w = get_variable("weights", [int(inputs.get_shape()[-1]), num_outputs],
                         initializer=weights_initializer or xavier_initializer(), regularizer=weights_regularizer,
                         trainable=trainable)
        out = matmul(inputs, w)
        if biases_initializer is not None:
            b = get_variable("biases", [num_outputs], initializer=biases_initializer, regularizer=biases_regularizer,
                             trainable=trainable)
            out = bias_add(out, b)
        return activation_fn(out) if activation_fn is not None else out


def flatten(x, *a, **k):
    v = _raw(x)
    return T(v.reshape(v.shape[0], -1))


# ---- tf.train ---------------------------------------------------------------------------------------
class MinimizeOp(object):
    def __init__(self, opt, loss, var_list):
        self.opt, self.loss, self.var_list = opt, loss, list(var_list)

    def gradients(self):
        gs = torch.autograd.grad(self.loss.v, [v.v for v in self.var_list], allow_unused=True, retain_graph=True)
        return {v.name[:-2]: (np.zeros(v.shape) if g is None else g.numpy().copy())
                for v, g in zip(self.var_list, gs)}


class AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss, var_list=None, **k):
        op = MinimizeOp(self, loss, var_list if var_list is not None else get_collection(GraphKeys.TRAINABLE_VARIABLES))
        STATE.minimize_ops.append(op)
        return op


class ExponentialMovingAverage(object):
    def __init__(self, decay, **k):
        self.decay = decay

    def apply(self, var_list=None):
        return ("ema_apply", list(var_list or []))

    def average(self, x):
        return x


# ---- module tree -----------------------------------------------------------------------------------
class _Missing(object):
    """Anything the reference merely *mentions* at import time (default arguments of the GAN / conv
    classes).  Calling it means the generator strayed outside the pinned slice."""

    def __init__(self, path):
        self._path = path

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Missing(self._path + "." + name)

    def __call__(self, *a, **k):
        raise NotImplementedError("tf_shim: %s is outside the shimmed slice" % self._path)


def _module(path, **attrs):
    mod = types.ModuleType(path)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda name, _p=path: _missing_attr(_p, name)
    sys.modules[path] = mod
    if "." in path:
        parent, leaf = path.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mod)
    return mod


def _missing_attr(path, name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing(path + "." + name)


def install():
    """Put the shim (and stubs for the two non-TF imports the reference needs) into sys.modules."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "_is_cfl_shim", False):
        return sys.modules["tensorflow"]
    g = globals()
    top = {k: g[k] for k in (
        "float32 float64 int32 int64 GraphKeys variable_scope name_scope get_variable get_collection "
        "add_to_collection reset_default_graph zeros_initializer ones_initializer constant_initializer reduce_sum "
        "reduce_mean reshape square sqrt subtract multiply matmul concat stack unstack split expand_dims transpose "
        "tile ones_like zeros_like constant greater less_equal cast group set_random_seed add_n maximum minimum "
        "clip_by_value gather_nd random_uniform placeholder placeholder_with_default").split()}
    tf = _module("tensorflow", _is_cfl_shim=True, range=range_, bool=bool_, **top)
    nn = dict(relu=relu, tanh=tanh, sigmoid=sigmoid, softmax=softmax, bias_add=bias_add, l2_loss=l2_loss,
              sigmoid_cross_entropy_with_logits=sigmoid_cross_entropy_with_logits)
    _module("tensorflow.nn", **nn)
    _module("tensorflow.train", AdamOptimizer=AdamOptimizer, ExponentialMovingAverage=ExponentialMovingAverage)
    noop = lambda *a, **k: None
    _module("tensorflow.summary", scalar=noop, image=noop, histogram=noop, merge=noop, FileWriter=noop)
    _module("tensorflow.image")
    _module("tensorflow.contrib")
    _module("tensorflow.contrib.layers", xavier_initializer=xavier_initializer, l2_regularizer=l2_regularizer,
            fully_connected=fully_connected, flatten=flatten)
    _module("tensorflow.contrib.framework")
    _module("tensorflow.contrib.framework.python")
    _module("tensorflow.contrib.framework.python.ops", add_arg_scope=lambda f: f)
    _module("tensorflow.contrib.layers.python")
    _module("tensorflow.contrib.layers.python.layers")
    _module("tensorflow.contrib.layers.python.layers.initializers", xavier_initializer=xavier_initializer)
    _module("tensorflow.contrib.layers.python.layers.utils", get_variable_collections=lambda s, n: None,
            collect_named_outputs=lambda collections, name, outputs: outputs)
    _module("tensorflow.python")
    _module("tensorflow.python.framework")
    _module("tensorflow.python.framework.ops", convert_to_tensor=lambda x, *a, **k: x if isinstance(x, T) else T(x),
            get_collection=get_collection, add_to_collection=add_to_collection)
    _module("tensorflow.python.ops")
    _module("tensorflow.python.ops.array_ops")
    _module("tensorflow.python.ops.init_ops", zeros_initializer=zeros_initializer)
    _module("tensorflow.python.ops.nn", **nn)
    _module("tensorflow.python.ops.variable_scope", variable_scope=variable_scope, get_variable=get_variable)
    _module("tensorflow.python.ops.variables", PartitionedVariable=type("PartitionedVariable", (), {}))
    _module("tensorflow.examples")
    _module("tensorflow.examples.tutorials")
    _module("tensorflow.examples.tutorials.mnist")
    # scipy.misc lost imread / imsave long ago; the reference imports them at module level only
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import scipy.misc as misc
    for name in ("imread", "imsave"):
        if not hasattr(misc, name):
            setattr(misc, name, _Missing("scipy.misc." + name))
    return tf
