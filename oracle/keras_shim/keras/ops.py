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


def zeros(shape, dtype=None):
    return torch.zeros([int(s) for s in shape], dtype=torch.get_default_dtype())


def cast(x, dtype):
    x = torch.as_tensor(x)
    return x.to(getattr(torch, dtype) if isinstance(dtype, str) else dtype)


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
