"""Weight import: NVIDIA-WaveGlow (PyTorch) state_dict -> the Keras-layout weight file the engine loads.

Mirrors what the reference does when it converts the torch.hub checkpoint: Conv1d `[out, in, k]` and
ConvTranspose1d `[in, out, k]` both become Keras `[k, in, out]` / `[k, out, in]` by the `[2, 1, 0]` transpose
(models/weights_converter.py:252-271), weight-norm is removed first (architectures/waveglow_arch.py:327-335
calls `remove_weightnorm`), and the fused `WN.k.cond_layer` of newer checkpoints is split per layer exactly as
`WaveglowBlock(fused=True)` slices it (waveglow_arch.py:119-121). No torch needed: values may be numpy arrays
or anything `np.asarray` accepts (call `.numpy()` / `.cpu()` on tensors first).
"""
from __future__ import annotations

import re

import numpy as np

from .weights import WaveGlowHParams, check_weights


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def _materialise_weight_norm(sd):
    """weight = g * v / ||v|| (norm over all dims but 0), as torch.nn.utils.remove_weight_norm does."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            v = _np(v)
            g = _np(sd[base + ".weight_g"])
            norm = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(axis=1)).reshape((-1,) + (1,) * (v.ndim - 1))
            out[base + ".weight"] = g.reshape(norm.shape) * v / norm
        elif k.endswith(".weight_g"):
            continue
        else:
            out[k] = _np(v)
    return out


def from_nvidia_state_dict(state_dict, *, n_early_every=4, n_early_size=2):
    """Returns (WaveGlowHParams, {keras name: float32 array})."""
    sd = _materialise_weight_norm(state_dict)
    t = lambda w: np.ascontiguousarray(np.transpose(w, (2, 1, 0)))      # torch conv layouts -> keras
    n_flows = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"convinv\.(\d+)\.", k)] if m)
    n_layers = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"WN\.0\.in_layers\.(\d+)\.", k)] if m)
    C = sd["WN.0.start.weight"].shape[0]
    n_group = sd["convinv.0.weight"].shape[0]
    n_mel = sd["upsample.weight"].shape[0]
    ks = sd["WN.0.in_layers.0.weight"].shape[2]
    hp = WaveGlowHParams(n_mel_channels=n_mel, n_flows=n_flows, n_group=n_group, n_early_every=n_early_every,
                         n_early_size=n_early_size, n_layers=n_layers, n_channels=C, kernel_size=ks)
    w = {"upsample/kernel": t(sd["upsample.weight"]), "upsample/bias": sd["upsample.bias"]}
    for k in range(n_flows):
        p, q = f"block-{k}/", f"WN.{k}."
        w[f"invertible_conv-{k}/conv/kernel"] = t(sd[f"convinv.{k}.weight"])
        w[p + "start_conv/kernel"], w[p + "start_conv/bias"] = t(sd[q + "start.weight"]), sd[q + "start.bias"]
        w[p + "end_conv/kernel"], w[p + "end_conv/bias"] = t(sd[q + "end.weight"]), sd[q + "end.bias"]
        fused = q + "cond_layer.weight" in sd
        for i in range(n_layers):
            w[p + f"in_conv-{i}/kernel"] = t(sd[q + f"in_layers.{i}.weight"])
            w[p + f"in_conv-{i}/bias"] = sd[q + f"in_layers.{i}.bias"]
            if fused:
                sl = slice(i * 2 * C, (i + 1) * 2 * C)
                w[p + f"cond_layer-{i}/kernel"] = t(sd[q + "cond_layer.weight"][sl])
                w[p + f"cond_layer-{i}/bias"] = sd[q + "cond_layer.bias"][sl]
            else:
                w[p + f"cond_layer-{i}/kernel"] = t(sd[q + f"cond_layers.{i}.weight"])
                w[p + f"cond_layer-{i}/bias"] = sd[q + f"cond_layers.{i}.bias"]
            w[p + f"res_skip_conv-{i}/kernel"] = t(sd[q + f"res_skip_layers.{i}.weight"])
            w[p + f"res_skip_conv-{i}/bias"] = sd[q + f"res_skip_layers.{i}.bias"]
    w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}
    check_weights(hp, w)
    return hp, w


def to_nvidia_state_dict(hp: WaveGlowHParams, weights, *, fused_cond=False):
    """Inverse mapping (used by the tests and to hand weights to NVIDIA's torch implementation)."""
    t = lambda w: np.ascontiguousarray(np.transpose(np.asarray(w, dtype=np.float32), (2, 1, 0)))
    sd = {"upsample.weight": t(weights["upsample/kernel"]), "upsample.bias": np.asarray(weights["upsample/bias"])}
    for k in range(hp.n_flows):
        p, q = f"block-{k}/", f"WN.{k}."
        sd[f"convinv.{k}.weight"] = t(weights[f"invertible_conv-{k}/conv/kernel"])
        sd[q + "start.weight"], sd[q + "start.bias"] = t(weights[p + "start_conv/kernel"]), weights[p + "start_conv/bias"]
        sd[q + "end.weight"], sd[q + "end.bias"] = t(weights[p + "end_conv/kernel"]), weights[p + "end_conv/bias"]
        cw, cb = [], []
        for i in range(hp.n_layers):
            sd[q + f"in_layers.{i}.weight"] = t(weights[p + f"in_conv-{i}/kernel"])
            sd[q + f"in_layers.{i}.bias"] = weights[p + f"in_conv-{i}/bias"]
            sd[q + f"res_skip_layers.{i}.weight"] = t(weights[p + f"res_skip_conv-{i}/kernel"])
            sd[q + f"res_skip_layers.{i}.bias"] = weights[p + f"res_skip_conv-{i}/bias"]
            if fused_cond:
                cw.append(t(weights[p + f"cond_layer-{i}/kernel"]))
                cb.append(np.asarray(weights[p + f"cond_layer-{i}/bias"]))
            else:
                sd[q + f"cond_layers.{i}.weight"] = t(weights[p + f"cond_layer-{i}/kernel"])
                sd[q + f"cond_layers.{i}.bias"] = weights[p + f"cond_layer-{i}/bias"]
        if fused_cond:
            sd[q + "cond_layer.weight"] = np.concatenate(cw, axis=0)
            sd[q + "cond_layer.bias"] = np.concatenate(cb, axis=0)
    return {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}


# ---- Keras 3 `.weights.h5` (models restored by `CheckpointManager.load`, custom_train_objects/checkpoint_manager.py:169-216)
_SUFFIX = re.compile(r"^(.*?)(?:_(\d+))?$")


def _list_index(name):
    """Keras names the saveables of a list attribute by class: 'conv1d', 'conv1d_1', 'conv1d_2', ... (the HDF5 library
    then stores the links alphabetically, so 'conv1d_10' sorts before 'conv1d_2'): the numeric suffix is the position."""
    m = _SUFFIX.match(name)
    return int(m.group(2)) if m.group(2) is not None else 0


def from_keras_weights_h5(path, *, n_early_every=4, n_early_size=2):
    """Reads the variables of the reference's `architectures.WaveGlow` from a Keras 3 `.weights.h5` file.

    Keras 3 (`keras/src/saving/saving_lib.py`) walks the model by ATTRIBUTE name and stores each layer's variables as
    `<path>/vars/<i>` in creation order (kernel 0, bias 1); list attributes get one sub-group per element named after
    the element's class in snake case with a running suffix. For the reference's model (waveglow_arch.py:164-223) that
    gives
        upsample/vars/{0,1}
        blocks/waveglow_block[_k]/{start,end}/vars/{0,1}
        blocks/waveglow_block[_k]/{in_layers,cond_layers,res_skip_layers}/conv1d[_i]/vars/{0,1}   (fused: cond_layer/vars)
        convinv/invertible1x1_conv[_k]/conv/vars/0
    A file whose groups carry the LAYER names instead (`block-3/in_conv-5/vars/0`, as older savers wrote) is accepted
    too. No Keras is installable here, so this layout is taken from the library's published saving algorithm and is
    NOT verified against a file written by Keras itself; the HDF5 container parsing is (text_to_speech_b200/h5lite.py).
    Returns (WaveGlowHParams, {keras variable name: float32 array}) like `from_nvidia_state_dict`."""
    from .h5lite import read_h5_datasets
    ds = {k.strip("/"): v for k, v in read_h5_datasets(path).items()}
    tree = {}
    for k, v in ds.items():
        parts = k.split("/")
        if len(parts) < 3 or parts[-2] != "vars":
            continue
        tree.setdefault(tuple(parts[:-2]), {})[int(parts[-1])] = np.asarray(v, dtype=np.float32)
    if not tree:
        raise ValueError(f"{path}: no '<layer>/vars/<i>' datasets found (not a Keras weights file?)")

    w = {}
    by_layer_name = any(p and re.fullmatch(r"block-\d+", p[0]) for p in tree)
    if by_layer_name:
        for p, vs in tree.items():
            base = "/".join(p)
            w[base + "/kernel"] = vs[0]
            if 1 in vs:
                w[base + "/bias"] = vs[1]
    else:
        def children(prefix):
            names = sorted({p[len(prefix)] for p in tree if len(p) > len(prefix) and p[:len(prefix)] == prefix}, key=_list_index)
            return names

        def put(dst, path_):
            if path_ not in tree:
                raise ValueError(f"{path}: group '{'/'.join(path_)}/vars' is missing")
            w[dst + "/kernel"] = tree[path_][0]
            if 1 in tree[path_]:
                w[dst + "/bias"] = tree[path_][1]

        put("upsample", ("upsample",))
        blocks, convs = children(("blocks",)), children(("convinv",))
        if not blocks or len(blocks) != len(convs):
            raise ValueError(f"{path}: found {len(blocks)} 'blocks' and {len(convs)} 'convinv' groups")
        for k, (bn, cn) in enumerate(zip(blocks, convs)):
            put(f"invertible_conv-{k}/conv", ("convinv", cn, "conv"))
            put(f"block-{k}/start_conv", ("blocks", bn, "start"))
            put(f"block-{k}/end_conv", ("blocks", bn, "end"))
            ins = children(("blocks", bn, "in_layers"))
            for i, name in enumerate(ins):
                put(f"block-{k}/in_conv-{i}", ("blocks", bn, "in_layers", name))
            for i, name in enumerate(children(("blocks", bn, "res_skip_layers"))):
                put(f"block-{k}/res_skip_conv-{i}", ("blocks", bn, "res_skip_layers", name))
            conds = children(("blocks", bn, "cond_layers"))
            if conds:
                for i, name in enumerate(conds):
                    put(f"block-{k}/cond_layer-{i}", ("blocks", bn, "cond_layers", name))
            else:                                               # fused=True: one 640 -> 2C*n_layers conv, sliced per layer
                put(f"block-{k}/__fused", ("blocks", bn, "cond_layer"))
                fk, fb = w.pop(f"block-{k}/__fused/kernel"), w.pop(f"block-{k}/__fused/bias")
                n_layers, C2 = len(ins), fk.shape[2] // len(ins)
                for i in range(n_layers):
                    w[f"block-{k}/cond_layer-{i}/kernel"] = np.ascontiguousarray(fk[:, :, i * C2:(i + 1) * C2])
                    w[f"block-{k}/cond_layer-{i}/bias"] = np.ascontiguousarray(fb[i * C2:(i + 1) * C2])
    n_flows = 1 + max(int(m.group(1)) for k in w for m in [re.match(r"block-(\d+)/", k)] if m)
    n_layers = 1 + max(int(m.group(1)) for k in w for m in [re.match(r"block-0/in_conv-(\d+)/", k)] if m)
    up = w["upsample/kernel"]
    hp = WaveGlowHParams(n_mel_channels=int(up.shape[1]), n_flows=n_flows,
                         n_group=int(w["invertible_conv-0/conv/kernel"].shape[1]), n_early_every=n_early_every,
                         n_early_size=n_early_size, n_layers=n_layers,
                         n_channels=int(w["block-0/start_conv/kernel"].shape[2]),
                         kernel_size=int(w["block-0/in_conv-0/kernel"].shape[0]))
    w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}
    check_weights(hp, w)
    return hp, w
