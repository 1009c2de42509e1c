"""keras.ops subset used by the reference WaveGlow (torch-backed)."""
import torch
import torch.nn.functional as F


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.get_default_dtype())


def tanh(x): return torch.tanh(_t(x))
def sigmoid(x): return torch.sigmoid(_t(x))
def exp(x): return torch.exp(_t(x))
def log(x): return torch.log(_t(x))
def det(x): return torch.linalg.det(_t(x))
def inv(x): return torch.linalg.inv(_t(x))
def shape(x): return tuple(x.shape)


def reshape(x, newshape):
    return _t(x).reshape([int(s) for s in newshape])


def transpose(x, axes=None):
    x = _t(x)
    if axes is None:
        axes = list(reversed(range(x.dim())))
    return x.permute(*axes)


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def _dtype(dtype):
    """Keras dtype names -> torch; every float maps to the default float dtype so a float64 run stays float64."""
    if dtype is None or (isinstance(dtype, str) and dtype.startswith("float")):
        return torch.get_default_dtype()
    if isinstance(dtype, torch.dtype):
        return torch.get_default_dtype() if dtype.is_floating_point else dtype
    return getattr(torch, dtype)


def zeros(shape, dtype=None):
    return torch.zeros([int(s) for s in shape], dtype=_dtype(dtype))


def cast(x, dtype):
    return torch.as_tensor(x).to(_dtype(dtype))


def pad(x, pad_width, mode="constant", constant_values=0):
    x = _t(x)
    flat = []
    for lo, hi in reversed(list(pad_width)):
        flat += [int(lo), int(hi)]
    return F.pad(x, flat, mode="constant", value=constant_values)


def conv(inputs, kernel, strides=1, padding="valid", data_format=None, dilation_rate=1):
    """keras.ops.conv, 1-D channels-last: inputs [B, L, in], kernel [k, in, out]."""
    x = _t(inputs).permute(0, 2, 1)
    w = _t(kernel).permute(2, 1, 0)
    k = w.shape[-1]
    if padding == "same":
        total = dilation_rate * (k - 1)
        x = F.pad(x, (total // 2, total - total // 2))
    elif padding != "valid":
        raise NotImplementedError(padding)
    y = F.conv1d(x, w, None, stride=strides, dilation=dilation_rate)
    return y.permute(0, 2, 1)


# ---- symbols the reference's Tacotron2 decoder / attention touch ----------------------------------------
def convert_to_tensor(x, dtype=None):
    return torch.as_tensor(x, dtype=_dtype(dtype))


def where(cond, x, y):
    x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.get_default_dtype())
    y = y if isinstance(y, torch.Tensor) else torch.as_tensor(y, dtype=x.dtype)
    return torch.where(cond, x, y)


def softmax(x, axis=-1): return torch.softmax(_t(x), dim=axis)
def matmul(a, b): return torch.matmul(_t(a), _t(b))
def stack(xs, axis=0): return torch.stack([_t(x) for x in xs], dim=axis)
def arange(start, stop=None, step=1, dtype=None):
    return torch.arange(start, stop, step, dtype=_dtype(dtype or "int32")) if stop is not None else \
        torch.arange(start, dtype=_dtype(dtype or "int32"))
def count_nonzero(x, axis=None): return torch.count_nonzero(torch.as_tensor(x), dim=axis).to(torch.int32)
def logical_and(a, b): return torch.logical_and(torch.as_tensor(a), torch.as_tensor(b))
def logical_or(a, b): return torch.logical_or(torch.as_tensor(a), torch.as_tensor(b))
def logical_not(a): return torch.logical_not(torch.as_tensor(a))
def argmax(x, axis=None): return torch.argmax(_t(x), dim=axis).to(torch.int32)
def all(x, axis=None): return torch.all(torch.as_tensor(x)) if axis is None else torch.all(torch.as_tensor(x), dim=axis)
def any(x, axis=None): return torch.any(torch.as_tensor(x)) if axis is None else torch.any(torch.as_tensor(x), dim=axis)
def maximum(a, b): return torch.maximum(torch.as_tensor(a), torch.as_tensor(b))
def minimum(a, b): return torch.minimum(torch.as_tensor(a), torch.as_tensor(b))
def tile(x, reps): return _t(x).repeat(*[int(r) for r in reps])
def eye(n, dtype=None): return torch.eye(int(n), dtype=_dtype(dtype))
def ones(shape, dtype=None): return torch.ones([int(s) for s in shape], dtype=_dtype(dtype))


def slice_update(inputs, start_indices, updates):
    out = inputs.clone()
    idx = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start_indices, updates.shape))
    out[idx] = updates.to(out.dtype)
    return out


def while_loop(cond, body, loop_vars, maximum_iterations=None):
    it = 0
    while (maximum_iterations is None or it < int(maximum_iterations)) and bool(cond(*loop_vars)):
        loop_vars = body(*loop_vars)
        it += 1
    return loop_vars
