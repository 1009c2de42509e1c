"""CPU: pins the oracle (restatement of architectures/waveglow_arch.py:244-306) against the committed
golden fixtures (produced by the reference's own source, oracle/gen_golden.py), against an independent
conv-op formulation, against the reference source itself when /root/reference is present, and against
algebraic properties that need no reference."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle.waveglow_oracle import OracleWaveGlow, infer_conv_ops, w_inverse
from oracle import run_reference
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights, synthetic_inputs

FAST_CASES = [c for c in GOLDEN_CASES if c != "wg256_k1"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_golden(case):
    hp, w, f = load_golden(case)
    sigma = float(f["sigma"])
    out = OracleWaveGlow(hp, w)(f["mel"], f["z"], sigma).numpy()
    assert out.shape == f["wave_reference_fp32"].shape
    assert np.abs(out - f["wave_reference_fp32"]).max() <= 1e-5          # reference source, fp32
    assert np.abs(out - f["wave_oracle_fp64"]).max() <= 1e-5             # fp64 arbiter
    det = OracleWaveGlow(hp, w)(f["mel"], None, sigma, deterministic=True).numpy()
    assert np.abs(det - f["wave_reference_deterministic"]).max() <= 1e-5


@pytest.mark.parametrize("case", ["tiny_c16", "nvidia_c32", "wg256_t24"])
def test_oracle_fp64_matches_golden(case):
    hp, w, f = load_golden(case)
    out = OracleWaveGlow(hp, w, torch.float64)(f["mel"], f["z"], float(f["sigma"])).numpy()
    assert np.abs(out - f["wave_oracle_fp64"]).max() <= 1e-9


@pytest.mark.parametrize("case", ["tiny_c16", "nvidia_c32", "wg256_t24"])
def test_independent_conv_formulation(case):
    hp, w, f = load_golden(case)
    a = OracleWaveGlow(hp, w)(f["mel"], f["z"], float(f["sigma"])).numpy()
    b = infer_conv_ops(hp, w, f["mel"], f["z"], float(f["sigma"])).numpy()
    assert np.abs(a - b).max() <= 1e-5


@pytest.mark.skipif(not run_reference.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("B,T,sigma,bias", [(1, 5, 1.0, 0.0), (3, 9, 0.6, 0.05)])
def test_restatement_equals_reference_source(B, T, sigma, bias):
    hp = WaveGlowHParams(n_channels=32)
    w = generate_weights(hp, 21, bias_std=bias)
    mel, z = synthetic_inputs(22, B, T, hp)
    ref = run_reference.reference_infer(hp, w, mel, z, sigma=sigma)
    out = OracleWaveGlow(hp, w)(mel, z, sigma).numpy()
    assert np.abs(ref - out).max() <= 1e-5


def test_zero_end_conv_makes_coupling_identity():
    # waveglow_arch.py:60-64: with the end conv at zero the affine coupling does nothing, so infer
    # reduces to the chain of W^-1 mixings and z re-injections -- independent of the mel.
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=2, n_channels=16)
    w = generate_weights(hp, 5, end_std=0.0)
    mel, z = synthetic_inputs(6, 2, 4, hp)
    mel2, _ = synthetic_inputs(7, 2, 4, hp)
    o = OracleWaveGlow(hp, w)
    a, b = o(mel, z, 0.8).numpy(), o(mel2, z, 0.8).numpy()
    assert np.abs(a - b).max() <= 1e-6
    # explicit chain
    zt = torch.as_tensor(z)
    n_rem = hp.n_remaining_channels
    audio, rest = 0.8 * zt[:, :, :n_rem], zt[:, :, n_rem:]
    for k in reversed(range(hp.n_flows)):
        audio = audio @ o.w_inv[k]
        if k % hp.n_early_every == 0 and k > 0:
            audio = torch.cat([0.8 * rest[:, :, :hp.n_early_size], audio], 2)
            rest = rest[:, :, hp.n_early_size:]
    assert np.abs(audio.reshape(2, -1).numpy() - a).max() <= 1e-5


def test_flow_is_invertible():
    # forward direction of the flows (invertible_conv.py:52-61 forward conv; coupling a1*exp(s)+b)
    # applied to the generated audio must give back sigma*z.
    hp = WaveGlowHParams(n_flows=4, n_early_every=2, n_layers=3, n_channels=16)
    w = generate_weights(hp, 9, bias_std=0.05)
    mel, z = synthetic_inputs(10, 1, 5, hp)
    o = OracleWaveGlow(hp, w, torch.float64)
    wave = o(mel, z, 0.7)
    spect = o.spect(torch.as_tensor(mel, dtype=torch.float64))
    audio = wave.reshape(1, -1, hp.n_group)
    outs = []
    for k in range(hp.n_flows):
        if k % hp.n_early_every == 0 and k > 0:
            outs.append(audio[:, :, :hp.n_early_size])
            audio = audio[:, :, hp.n_early_size:]
        kern = torch.as_tensor(w[f"invertible_conv-{k}/conv/kernel"], dtype=torch.float64)
        audio = audio @ kern[0]                                    # forward 1x1 conv
        nh = audio.shape[2] // 2
        a0, a1 = audio[:, :, :nh], audio[:, :, nh:]
        out = o.wn_block(k, a0, spect)
        audio = torch.cat([a0, a1 * torch.exp(out[:, :, nh:]) + out[:, :, :nh]], 2)
    outs.append(audio)
    zrec = torch.cat(list(reversed(outs)), 2) if False else None
    # z order: the first n_rem channels seed the last flow; early outputs were taken in front
    z_t = torch.as_tensor(z, dtype=torch.float64)
    n_rem = hp.n_remaining_channels
    assert torch.allclose(audio, 0.7 * z_t[:, :, :n_rem], atol=1e-6)
    # flows run 3,2,1,0 in infer and inject after flow 2 (k % 2 == 0, k > 0): that chunk is z[4:6],
    # recovered here at the top of the forward loop for k = 2
    assert len(outs) == 2 and torch.allclose(outs[0], 0.7 * z_t[:, :, n_rem:n_rem + 2], atol=1e-6)


def test_w_inverse_layout():
    k = torch.randn(1, 6, 6)
    M = w_inverse(k)
    x = torch.randn(3, 5, 6)
    y = x @ k[0]                    # forward conv1x1 with keras kernel [1,in,out]
    assert torch.allclose(y @ M, x, atol=1e-4)
