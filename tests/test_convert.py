"""CPU: NVIDIA-WaveGlow state_dict import. The converted weights must reproduce, through the
oracle, what NVIDIA's channels-first formulation computes from the torch-layout tensors."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.waveglow_oracle import OracleWaveGlow
from text_to_speech_b200.convert import from_nvidia_state_dict, to_nvidia_state_dict
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs, weights_digest


@pytest.mark.parametrize("fused", [False, True])
def test_state_dict_round_trip(fused):
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=3, n_channels=16)
    w = generate_weights(hp, 3, bias_std=0.1)
    sd = to_nvidia_state_dict(hp, w, fused_cond=fused)
    assert sd["upsample.weight"].shape == (80, 80, 1024) and sd["WN.0.in_layers.0.weight"].shape == (32, 16, 3)
    hp2, w2 = from_nvidia_state_dict(sd, n_early_every=2, n_early_size=2)
    assert hp2 == hp and weights_digest(w2) == weights_digest(w)


def test_weight_norm_is_removed():
    hp = WaveGlowHParams(n_flows=2, n_early_every=2, n_layers=2, n_channels=16)
    w = generate_weights(hp, 4)
    sd = to_nvidia_state_dict(hp, w)
    wt = sd.pop("WN.1.in_layers.1.weight")
    g = np.sqrt((wt.reshape(wt.shape[0], -1) ** 2).sum(1)).reshape(-1, 1, 1)
    sd["WN.1.in_layers.1.weight_g"], sd["WN.1.in_layers.1.weight_v"] = g, 3.0 * wt   # any positive rescale of v
    _, w2 = from_nvidia_state_dict(sd, n_early_every=2)
    assert np.allclose(w2["block-1/in_conv-1/kernel"], w["block-1/in_conv-1/kernel"], atol=1e-6)


def test_torch_layout_tensors_compute_the_same_function():
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=2, n_channels=16)
    w = generate_weights(hp, 5, bias_std=0.05)
    sd = {k: torch.from_numpy(v) for k, v in to_nvidia_state_dict(hp, w).items()}
    mel, z = synthetic_inputs(6, 1, 5, hp)
    # one WN input conv computed NVIDIA-style (channels first) vs the oracle's channels-last restatement
    x = torch.randn(1, 16, 40)
    y_t = F.conv1d(x, sd["WN.2.in_layers.1.weight"], sd["WN.2.in_layers.1.bias"], dilation=2, padding=2)
    from oracle.waveglow_oracle import dilated_conv
    _, w2 = from_nvidia_state_dict({k: v.numpy() for k, v in sd.items()}, n_early_every=2)
    y_k = dilated_conv(x.permute(0, 2, 1), torch.from_numpy(w2["block-2/in_conv-1/kernel"]),
                       torch.from_numpy(w2["block-2/in_conv-1/bias"]), 2)
    assert torch.allclose(y_t.permute(0, 2, 1), y_k, atol=1e-5)
    out = OracleWaveGlow(hp, w2)(mel, z, 0.6).numpy()
    assert np.isfinite(out).all() and out.shape == (1, 5 * 256)


REF_CONVERTER = "/root/reference/models/weights_converter.py"


@pytest.mark.skipif(not __import__("os").path.isfile(REF_CONVERTER), reason="reference tree not mounted")
def test_layout_rule_is_the_reference_transpose_weights():
    """The reference converts torch tensors to Keras layouts with `transpose_weights` (models/weights_converter.py:252-271,
    a stand-alone numpy function): every conv kernel `from_nvidia_state_dict` emits must be exactly what that REAL
    function returns for the NVIDIA tensor (conditioning convs: for the per-layer slice)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_weights_converter", REF_CONVERTER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=3, n_channels=16)
    w = generate_weights(hp, 9, bias_std=0.1)
    for fused in (False, True):
        sd = to_nvidia_state_dict(hp, w, fused_cond=fused)
        _, got = from_nvidia_state_dict(sd, n_early_every=2, n_early_size=2)
        checked = 0
        pairs = [("upsample.weight", "upsample/kernel")]
        for k in range(hp.n_flows):
            pairs += [(f"WN.{k}.start.weight", f"block-{k}/start_conv/kernel"), (f"WN.{k}.end.weight", f"block-{k}/end_conv/kernel")]
            for i in range(hp.n_layers):
                pairs += [(f"WN.{k}.in_layers.{i}.weight", f"block-{k}/in_conv-{i}/kernel"),
                          (f"WN.{k}.res_skip_layers.{i}.weight", f"block-{k}/res_skip_conv-{i}/kernel")]
                if not fused:
                    pairs.append((f"WN.{k}.cond_layers.{i}.weight", f"block-{k}/cond_layer-{i}/kernel"))
        for src, dst in pairs:
            if src in sd and dst in got:
                assert np.array_equal(ref.transpose_weights(np.asarray(sd[src])), got[dst]), (src, dst)
                checked += 1
        assert checked >= (2 + 2 * hp.n_layers) * hp.n_flows
        if fused:       # one 640 -> 2C*n_layers conv per flow, sliced per layer after the reference's transpose
            C2 = 2 * hp.n_channels
            for k in range(hp.n_flows):
                full = ref.transpose_weights(np.asarray(sd[f"WN.{k}.cond_layer.weight"]))
                for i in range(hp.n_layers):
                    assert np.array_equal(full[..., i * C2:(i + 1) * C2], got[f"block-{k}/cond_layer-{i}/kernel"])


def test_h5_reader_on_a_file_written_by_the_real_hdf5_library():
    """The only genuine HDF5 file on this machine: scipy's MATLAB v7.3 test file (512-byte user block, written by the
    HDF5 library itself). scipy's own test expects `testdouble` = linspace(0, 2 pi, 9)."""
    import glob
    import scipy.io
    from text_to_speech_b200.h5lite import read_h5_datasets
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5*.mat"))
    if not hits:
        pytest.skip("scipy's HDF5 test file is not installed")
    d = read_h5_datasets(hits[0])
    assert list(d) == ["testdouble"] and d["testdouble"].dtype == np.float64
    assert np.allclose(d["testdouble"].ravel(), np.linspace(0, 2 * np.pi, 9))


@pytest.mark.parametrize("fused", [False, True])
def test_keras_weights_h5_import_round_trip(tmp_path, fused):
    """`.weights.h5` import (checkpoint_manager.py:169-216): a file laid out as Keras 3 lays out the reference's
    WaveGlow (attribute-named groups, list elements conv1d / conv1d_1 / ... stored ALPHABETICALLY by the HDF5 library,
    > 8 links per group = several symbol-table nodes) is read back to exactly the weight set and hparams."""
    from oracle.h5_writer import write_h5, keras3_waveglow_layout
    from text_to_speech_b200.convert import from_keras_weights_h5
    from text_to_speech_b200.h5lite import H5File, H5FormatError
    hp = WaveGlowHParams(n_flows=12, n_layers=11, n_channels=16)          # 11 layers: conv1d_10 sorts before conv1d_2
    w = generate_weights(hp, 3, bias_std=0.1)
    path = str(tmp_path / "waveglow.weights.h5")
    write_h5(path, keras3_waveglow_layout(hp, w, fused=fused))
    with H5File(path) as f:
        names = [p for p, _, g in f.walk() if not g]
    n_vars = len([k for k in w if not k.startswith("__")])
    assert len(names) == (n_vars if not fused else n_vars - 2 * hp.n_flows * (hp.n_layers - 1))
    assert names.index("blocks/waveglow_block/in_layers/conv1d_10/vars/0") < names.index("blocks/waveglow_block/in_layers/conv1d_2/vars/0")
    hp2, w2 = from_keras_weights_h5(path)
    assert hp2 == hp
    assert sorted(w2) == sorted(k for k in w if not k.startswith("__"))
    for k in w2:
        assert np.array_equal(w2[k], w[k]), k
    # layer-named groups (block-3/in_conv-5/vars/0) are accepted as well
    alt = {}
    for k, v in w.items():
        if not k.startswith("__"):
            alt[k.rsplit("/", 1)[0] + "/vars/" + ("0" if k.endswith("kernel") else "1")] = v
    write_h5(path, alt)
    hp3, w3 = from_keras_weights_h5(path)
    assert hp3 == hp and all(np.array_equal(w3[k], w[k]) for k in w3)
    with open(path, "wb") as f:
        f.write(b"not an hdf5 file at all")
    with pytest.raises(H5FormatError):
        from_keras_weights_h5(path)
