"""keras.layers subset used by the reference WaveGlow (torch-backed, inference only)."""
import math

import torch
import torch.nn.functional as F

from . import ops


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False

    def build(self, input_shape=None):
        self.built = True

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    # -- variable tracking: walk attributes (incl. lists) for sub-layers --------------------------
    def _sublayers(self):
        for v in vars(self).values():
            if isinstance(v, Layer):
                yield v
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, Layer):
                        yield e

    def _named_variables(self, prefix=""):
        me = prefix + (self.name or type(self).__name__)
        for attr in ("kernel", "bias"):
            if isinstance(getattr(self, attr, None), torch.Tensor):
                yield me + "/" + attr, self, attr
        for sub in self._sublayers():
            yield from sub._named_variables(me + "/")


class Model(Layer):
    def set_weights(self, weights, **kwargs):
        """Shim: ``weights`` is a dict {variable path below the model: array}."""
        root = (self.name or type(self).__name__) + "/"
        seen = set()
        for path, layer, attr in self._named_variables():
            key = path[len(root):]
            if key not in weights:
                raise KeyError(f"no value for variable {key}")
            cur = getattr(layer, attr)
            val = torch.as_tensor(weights[key], dtype=cur.dtype)
            if tuple(val.shape) != tuple(cur.shape):
                raise ValueError(f"{key}: shape {tuple(val.shape)} != {tuple(cur.shape)}")
            setattr(layer, attr, val.clone())
            seen.add(key)
        extra = {k for k in weights if not k.startswith("__")} - seen
        if extra:
            raise KeyError(f"unused weights: {sorted(extra)[:5]}")

    def named_variables(self):
        root = (self.name or type(self).__name__) + "/"
        return {path[len(root):]: getattr(layer, attr) for path, layer, attr in self._named_variables()}


def _init(initializer, shape):
    dtype = torch.get_default_dtype()
    if initializer == "zeros":
        return torch.zeros(shape, dtype=dtype)
    # glorot_uniform (Keras default)
    receptive = int(math.prod(shape[:-2])) if len(shape) > 2 else 1
    limit = math.sqrt(6.0 / (shape[-2] * receptive + shape[-1] * receptive))
    return (torch.rand(shape, dtype=dtype) * 2 - 1) * limit


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", dilation_rate=1,
                 use_bias=True, kernel_initializer="glorot_uniform", name=None, **kwargs):
        super().__init__(name=name)
        self.filters = filters
        self.kernel_size = (kernel_size,) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.strides = (strides,) if isinstance(strides, int) else tuple(strides)
        self.dilation_rate = (dilation_rate,) if isinstance(dilation_rate, int) else tuple(dilation_rate)
        self.padding = padding
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        super().build(input_shape)
        in_ch = int(input_shape[-1])
        self.kernel = _init(self.kernel_initializer, (self.kernel_size[0], in_ch, self.filters))
        if self.use_bias:
            self.bias = torch.zeros(self.filters, dtype=torch.get_default_dtype())

    @property
    def weights(self):
        return [self.kernel] + ([self.bias] if self.use_bias else [])

    def load_own_variables(self, store):
        self.kernel = torch.as_tensor(store["0"], dtype=self.kernel.dtype)
        if self.use_bias:
            self.bias = torch.as_tensor(store["1"], dtype=self.kernel.dtype)

    def call(self, inputs):
        if not self.built:
            self.build(inputs.shape)
        y = ops.conv(inputs, self.kernel, strides=self.strides[0], padding=self.padding,
                     dilation_rate=self.dilation_rate[0])
        return y + self.bias if self.use_bias else y


class Conv1DTranspose(Layer):
    """kernel [k, out(filters), in]; 'valid': out length = (T-1)*stride + k."""
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True,
                 kernel_initializer="glorot_uniform", name=None, **kwargs):
        super().__init__(name=name)
        if padding != "valid":
            raise NotImplementedError(padding)
        self.filters = filters
        self.kernel_size = (kernel_size,) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.strides = (strides,) if isinstance(strides, int) else tuple(strides)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        super().build(input_shape)
        in_ch = int(input_shape[-1])
        self.kernel = _init(self.kernel_initializer, (self.kernel_size[0], self.filters, in_ch))
        if self.use_bias:
            self.bias = torch.zeros(self.filters, dtype=torch.get_default_dtype())

    def call(self, inputs):
        x = inputs.permute(0, 2, 1)
        w = self.kernel.permute(2, 1, 0)        # [k,out,in] -> torch conv_transpose1d [in,out,k]
        y = F.conv_transpose1d(x, w, self.bias if self.use_bias else None, stride=self.strides[0])
        return y.permute(0, 2, 1)
