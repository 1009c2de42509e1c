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
        for attr in ("kernel", "recurrent_kernel", "bias"):
            if isinstance(getattr(self, attr, None), torch.Tensor):
                yield me + "/" + attr, self, attr
        for sub in self._sublayers():
            yield from sub._named_variables(me + "/")


    @property
    def _layers(self):          # what keras.Layer tracks; custom_rnn_dropout_cell.py:63 walks it
        return list(self._sublayers())


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


# ---- symbols the reference's Tacotron2 DECODER touches (tacotron2_arch.py:143-212, 336-750;
# ---- layers/location_sensitive_attention.py) -- Keras 3 semantics restated ----------------------------

_ACTIVATIONS = {None: lambda x: x, "linear": lambda x: x, "relu": torch.relu, "tanh": torch.tanh,
                "sigmoid": torch.sigmoid}


class Input:
    """Placeholder marker (keras.layers.Input); Sequential skips it."""
    def __init__(self, shape=None, dtype=None, name=None, **kwargs):
        self.shape, self.dtype, self.name = shape, dtype, name


class Dense(Layer):
    """kernel [in, units]; y = activation(x @ kernel + bias)."""
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.use_bias, self.kernel_initializer = int(units), use_bias, kernel_initializer
        self.activation = _ACTIVATIONS[activation]
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        super().build(input_shape)
        init = self.kernel_initializer if isinstance(self.kernel_initializer, str) else "glorot_uniform"
        self.kernel = _init(init, (int(input_shape[-1]), self.units))
        if self.use_bias:
            self.bias = torch.zeros(self.units, dtype=torch.get_default_dtype())

    def call(self, inputs):
        if not self.built:
            self.build(inputs.shape)
        y = inputs @ self.kernel
        return self.activation(y + self.bias if self.use_bias else y)


class LSTMCell(Layer):
    """keras.layers.LSTMCell: kernel [in, 4u], recurrent_kernel [u, 4u], bias [4u] (unit forget bias), gate order
    i, f, c, o; activation tanh, recurrent activation sigmoid. cell(x, [h, c]) -> (h', [h', c'])."""
    def __init__(self, units, dropout=0.0, recurrent_dropout=0.0, name=None, **kwargs):
        super().__init__(name=name)
        self.units = int(units)
        self.kernel = self.recurrent_kernel = self.bias = None

    def build(self, input_shape):
        super().build(input_shape)
        u = self.units
        self.kernel = _init("glorot_uniform", (int(input_shape[-1]), 4 * u))
        self.recurrent_kernel = _init("glorot_uniform", (u, 4 * u))
        self.bias = torch.zeros(4 * u, dtype=torch.get_default_dtype())
        self.bias[u:2 * u] = 1.0

    def get_initial_state(self, batch_size=None):
        z = torch.zeros(int(batch_size), self.units, dtype=torch.get_default_dtype())
        return [z, z.clone()]

    def call(self, inputs, states, training=False):
        h, c = states
        z = inputs @ self.kernel + h @ self.recurrent_kernel + self.bias
        i, f, g, o = torch.chunk(z, 4, dim=-1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        return h_new, [h_new, c_new]

    # dropout-mask bookkeeping the reference pokes at (tacotron2_arch.py:374-393); rates are 0 at inference
    def get_dropout_mask(self, *a, **k): return None
    def get_recurrent_dropout_mask(self, *a, **k): return None
    def reset_dropout_mask(self): pass
    def reset_recurrent_dropout_mask(self): pass


class StackedRNNCells(Layer):
    """keras.layers.StackedRNNCells: states are a list with one entry per cell; with a single cell `call` returns
    that cell's state un-nested (keras/src/layers/rnn/stacked_rnn_cells.py)."""
    def __init__(self, cells, name=None, **kwargs):
        super().__init__(name=name)
        self.cells = list(cells)

    def build(self, input_shape):
        super().build(input_shape)
        shape = tuple(input_shape)
        for cell in self.cells:
            cell.build(shape)
            shape = shape[:-1] + (cell.units,)

    def get_initial_state(self, batch_size=None):
        return [cell.get_initial_state(batch_size=batch_size) for cell in self.cells]

    def call(self, inputs, states, training=False):
        new_states = []
        for cell, st in zip(self.cells, states):
            inputs, st = cell(inputs, list(st))
            new_states.append(st)
        if len(new_states) == 1:
            new_states = new_states[0]
        return inputs, new_states


class Sequential(Layer):
    def __init__(self, layers=None, name=None, **kwargs):
        super().__init__(name=name)
        self.layers = [l for l in (layers or []) if not isinstance(l, Input)]
        first = (layers or [None])[0]
        if isinstance(first, Input) and first.shape is not None:     # keras builds a Sequential that starts with an Input
            shape = (None,) + tuple(first.shape)
            for layer in self.layers:
                layer.build(shape)
                shape = shape[:-1] + (getattr(layer, "filters", None) or layer.units,)
            self.built = True

    def call(self, inputs, **kwargs):
        x = inputs
        for layer in self.layers:
            x = layer(x)
        return x
