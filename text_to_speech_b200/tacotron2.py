"""The CALLER of the WaveGlow path: a Tacotron2 mel producer for the end-to-end `tts()` workload
(BASELINE.json configs[3]).

STATUS -- read before citing a number from this file: this is NOT part of the vocoder's drop-in
boundary. It is a restatement of the reference's Tacotron2 inference (`architectures/tacotron2_arch.py
:609-749, 866-925`; attention `architectures/layers/location_sensitive_attention.py:96-186`) in plain torch
ops (cuDNN/cuBLAS underneath), written so that the vocoder stage can be measured in the pipeline it lives
in, plus the host side of the CUDA decoder loop (`decode_b200` -> csrc/taco.cu). Parity:
  * DECODER (prenet, attention LSTM, location sensitive attention, decoder LSTM, projection, stop gate, the
    finished/lengths bookkeeping): pinned to the reference's own source, executed unmodified over the Keras
    shim (run_reference_taco in the test oracle) -- `decode` reproduces it to 2e-16 in float64 (tests/test_oracle_taco.py,
    fixtures tests/golden/taco_decoder_*.npz);
  * ENCODER and POSTNET: the reference builds them with its generic `simple_cnn` factory on the functional
    Keras API, which the shim does not cover -- UNPINNED; tests cover internal consistency only (float64
    twin, padding invariance).

What is restated (inference only, single speaker):
  encoder   embedding(148, 512, pad 0) -> 3 x [conv k5 'same' -> batch-norm -> relu] -> BiLSTM(256+256),
            padded positions masked                                 (tacotron2_arch.py:235-333)
  decoder   per frame: prenet 80->256->256 (relu, dropout 0.5 kept ON at inference unless
            deterministic, :188-203) -> attention LSTM(1024) on [prenet, context] -> location
            sensitive attention (query 1024->128, memory 512->128, location conv k31 over
            [weights, cumulative weights] -> 32 -> 128, v . tanh(.), masked softmax) -> decoder
            LSTM(1024) on [attention h, context] -> [decoder h, context] -> frame (80) and stop
            probability (sigmoid)                                    (:422-486, :640-691)
  loop      zero first frame, `finished |= stop > 0.5`, `lengths += !finished`, optional early stop,
            mask = arange <= lengths                                 (:693-749)
  postnet   5 x [conv k5 'same' -> batch-norm -> tanh (none on the last)], mel = decoder + postnet (:214-233, :917-919)

The decode loop is launch-bound in eager mode (~35 small kernels per frame); `use_graph=True`
captures `graph_chunk` consecutive steps into one CUDA graph and replays it.
"""
from __future__ import annotations

import collections
import dataclasses
import math

import numpy as np
import torch
import torch.nn.functional as F

Tacotron2InferenceOutput = collections.namedtuple(
    "Tacotron2InferenceOutput", ["decoder_output", "mel", "stop_tokens", "attention_weights", "lengths"])


@dataclasses.dataclass(frozen=True)
class Tacotron2HParams:
    """Defaults of HParamsTacotron2 (tacotron2_arch.py:59-135) and HParamsLSA (location_sensitive_attention.py:17-24)."""
    vocab_size: int = 148
    pad_token: int = 0
    embedding_dim: int = 512
    encoder_n_conv: int = 3
    encoder_kernel_size: int = 5
    prenet_sizes: tuple = (256, 256)
    prenet_drop_rate: float = 0.5
    attention_rnn_dim: int = 1024
    decoder_rnn_dim: int = 1024
    attention_dim: int = 128
    attention_filters: int = 32
    attention_kernel_size: int = 31
    n_mel_channels: int = 80
    postnet_n_conv: int = 5
    postnet_filters: int = 512
    postnet_kernel_size: int = 5
    bn_epsilon: float = 1e-5


def generate_tacotron2_weights(hp: Tacotron2HParams, seed: int) -> dict:
    """Seeded random-init weights in KERAS layouts (Dense [in, out]; Conv1D [k, in, out]; LSTM kernel
    [in, 4u] / recurrent_kernel [u, 4u] / bias [4u] in gate order i, f, c, o with unit forget bias)."""
    rng = np.random.default_rng(seed)
    w = {}

    def glorot(name, *shape):
        fan_in = int(np.prod(shape[:-1]))
        fan_out = shape[-1] * (shape[0] if len(shape) == 3 else 1)
        if len(shape) == 3:
            fan_in = shape[0] * shape[1]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)

    def lstm(prefix, n_in, units):
        glorot(prefix + "/kernel", n_in, 4 * units)
        w[prefix + "/recurrent_kernel"] = (rng.standard_normal((units, 4 * units)) / math.sqrt(units)).astype(np.float32)
        b = np.zeros(4 * units, np.float32)
        b[units:2 * units] = 1.0
        w[prefix + "/bias"] = b

    def bnorm(prefix, ch):
        w[prefix + "/gamma"] = (1.0 + 0.1 * rng.standard_normal(ch)).astype(np.float32)
        w[prefix + "/beta"] = (0.1 * rng.standard_normal(ch)).astype(np.float32)
        w[prefix + "/moving_mean"] = (0.1 * rng.standard_normal(ch)).astype(np.float32)
        w[prefix + "/moving_variance"] = rng.uniform(0.5, 1.5, ch).astype(np.float32)

    E = hp.embedding_dim
    w["encoder/embeddings"] = rng.uniform(-0.05, 0.05, (hp.vocab_size, E)).astype(np.float32)
    for i in range(hp.encoder_n_conv):
        glorot(f"encoder/conv_{i}/kernel", hp.encoder_kernel_size, E, E)
        w[f"encoder/conv_{i}/bias"] = np.zeros(E, np.float32)
        bnorm(f"encoder/bn_{i}", E)
    lstm("encoder/bi_lstm/forward", E, E // 2)
    lstm("encoder/bi_lstm/backward", E, E // 2)
    n_in = hp.n_mel_channels
    for i, size in enumerate(hp.prenet_sizes):
        glorot(f"decoder/prenet/layer_{i}/kernel", n_in, size)
        n_in = size
    lstm("decoder/attention_rnn", hp.prenet_sizes[-1] + E, hp.attention_rnn_dim)
    glorot("decoder/lsa/query_layer/kernel", hp.attention_rnn_dim, hp.attention_dim)
    glorot("decoder/lsa/memory_layer/kernel", E, hp.attention_dim)
    glorot("decoder/lsa/value_layer/kernel", hp.attention_dim, 1)
    glorot("decoder/lsa/location_conv/kernel", hp.attention_kernel_size, 2, hp.attention_filters)
    glorot("decoder/lsa/location_dense/kernel", hp.attention_filters, hp.attention_dim)
    lstm("decoder/decoder_rnn/cell_0", hp.attention_rnn_dim + E, hp.decoder_rnn_dim)
    glorot("decoder/linear_projection/kernel", hp.decoder_rnn_dim + E, hp.n_mel_channels)
    w["decoder/linear_projection/bias"] = np.zeros(hp.n_mel_channels, np.float32)
    glorot("decoder/gate_output/kernel", hp.decoder_rnn_dim + E, 1)
    w["decoder/gate_output/bias"] = np.zeros(1, np.float32)
    ch = [hp.n_mel_channels] + [hp.postnet_filters] * (hp.postnet_n_conv - 1) + [hp.n_mel_channels]
    for i in range(hp.postnet_n_conv):
        glorot(f"postnet/conv_{i}/kernel", hp.postnet_kernel_size, ch[i], ch[i + 1])
        w[f"postnet/conv_{i}/bias"] = np.zeros(ch[i + 1], np.float32)
        bnorm(f"postnet/bn_{i}", ch[i + 1])
    return w


class Tacotron2:
    """`Tacotron2(hp, weights, device, dtype).infer(tokens, max_length=..., early_stopping=...)` -- the
    argument names and the returned namedtuple follow `architectures.Tacotron2.infer` (:866-925)."""

    def __init__(self, hp: Tacotron2HParams, weights: dict, device="cuda", dtype=torch.float32, b200_lstm_weights="split_bf16"):
        if b200_lstm_weights not in ("fp32", "bf16", "split_bf16"):
            raise ValueError("b200_lstm_weights must be 'fp32', 'bf16' or 'split_bf16'")
        self.hp, self.device, self.dtype, self.b200_lstm_weights = hp, torch.device(device), dtype, b200_lstm_weights
        t = lambda k: torch.as_tensor(np.asarray(weights[k]), dtype=dtype, device=self.device)  # noqa: E731
        self.emb = t("encoder/embeddings")
        self.enc_convs = []
        for i in range(hp.encoder_n_conv):
            self.enc_convs.append(self._conv_bn(t, f"encoder/conv_{i}", f"encoder/bn_{i}"))
        E = hp.embedding_dim
        self.bilstm = torch.nn.LSTM(E, E // 2, batch_first=True, bidirectional=True).to(self.device, dtype)
        with torch.no_grad():
            for sfx, d in (("", "forward"), ("_reverse", "backward")):
                getattr(self.bilstm, "weight_ih_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/kernel").T)
                getattr(self.bilstm, "weight_hh_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/recurrent_kernel").T)
                getattr(self.bilstm, "bias_ih_l0" + sfx).copy_(t(f"encoder/bi_lstm/{d}/bias"))
                getattr(self.bilstm, "bias_hh_l0" + sfx).zero_()
        self.bilstm.requires_grad_(False)
        self.prenet = [t(f"decoder/prenet/layer_{i}/kernel") for i in range(len(hp.prenet_sizes))]
        # LSTM cells as one fused [in + units, 4 units] matrix each
        self.att_w = torch.cat([t("decoder/attention_rnn/kernel"), t("decoder/attention_rnn/recurrent_kernel")])
        self.att_b = t("decoder/attention_rnn/bias")
        self.dec_w = torch.cat([t("decoder/decoder_rnn/cell_0/kernel"), t("decoder/decoder_rnn/cell_0/recurrent_kernel")])
        self.dec_b = t("decoder/decoder_rnn/cell_0/bias")
        self.q_w = t("decoder/lsa/query_layer/kernel")
        self.m_w = t("decoder/lsa/memory_layer/kernel")
        self.v_w = t("decoder/lsa/value_layer/kernel")[:, 0]
        self.loc_conv = t("decoder/lsa/location_conv/kernel").permute(2, 1, 0).contiguous()     # [filters, 2, k]
        self.loc_dense = t("decoder/lsa/location_dense/kernel")
        # frame projection and stop gate share their input: one [1536, 81] matrix
        self.out_w = torch.cat([t("decoder/linear_projection/kernel"), t("decoder/gate_output/kernel")], dim=1)
        self.out_b = torch.cat([t("decoder/linear_projection/bias"), t("decoder/gate_output/bias")])
        self.post_convs = [self._conv_bn(t, f"postnet/conv_{i}", f"postnet/bn_{i}") for i in range(hp.postnet_n_conv)]
        self._graphs = {}
        self.max_graphs = 8
        self._decoder_weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()
                                 if k.startswith("decoder/")}
        self._b200 = None

    def close(self):
        if self._b200 is not None:
            lib, h = self._b200
            lib.wg_taco_destroy(h)
            self._b200 = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- B200 decoder loop (include/wg_taco_b200.h)
    def _b200_handle(self):
        if self._b200 is None:
            import ctypes
            from . import _lib
            if self.device.type != "cuda" or self.dtype != torch.float32:
                raise RuntimeError("decoder='b200' needs a CUDA float32 model (there is no CPU fallback)")
            lib = _lib.load_library()
            hp = self.hp
            cfg = _lib.WgTacoConfig(hp.n_mel_channels, hp.prenet_sizes[-1], hp.embedding_dim, hp.attention_rnn_dim,
                                    hp.decoder_rnn_dim, hp.attention_dim, hp.attention_filters,
                                    hp.attention_kernel_size, hp.prenet_drop_rate,
                                    {"fp32": 0, "bf16": 1, "split_bf16": 2}[self.b200_lstm_weights])
            if tuple(hp.prenet_sizes) != (hp.prenet_sizes[-1],) * 2:
                raise RuntimeError("decoder='b200' supports a two-layer prenet of equal widths")
            names = sorted(self._decoder_weights)
            ts = (_lib.WgTensor * len(names))()
            for t, k in zip(ts, names):
                a = self._decoder_weights[k]
                t.name, t.ndim = k.encode(), a.ndim
                t.data = a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
                for i, n in enumerate(a.shape):
                    t.shape[i] = n
            h = ctypes.c_void_p()
            rc = lib.wg_taco_create(ctypes.byref(cfg), ts, len(names), self.device.index or 0, ctypes.byref(h))
            if rc != _lib.WG_OK:
                raise RuntimeError(f"wg_taco_create failed ({rc}): {lib.wg_taco_last_error(None).decode()}")
            self._b200 = (lib, h)
        return self._b200

    def decode_b200(self, memory, mask, max_length, early_stopping=True, deterministic=False, seed=0, graph_chunk=32,
                    return_attention=True):
        """The decoder loop on the B200 kernels (csrc/taco.cu): same arguments and results as `decode`."""
        import ctypes
        lib, h = self._b200_handle()
        B, S, _ = memory.shape
        text_len = mask.sum(1).to(torch.int32)
        if not bool((mask == (torch.arange(S, device=mask.device)[None] < text_len[:, None])).all()):
            raise ValueError("decoder='b200' needs right-padded texts (the mask must be a prefix mask)")
        tl = text_len.cpu().numpy().astype(np.int32)
        memory = memory.contiguous()
        z = lambda *sh, dt=torch.float32: torch.zeros(*sh, dtype=dt, device=self.device)  # noqa: E731
        outputs, stops = z(B, max_length, self.hp.n_mel_channels), z(B, max_length)
        attn = z(B, max_length, S) if return_attention else None
        lengths = z(B, dt=torch.int32)
        frames = ctypes.c_int32()
        lib.wg_taco_set_graph_chunk(h, int(graph_chunk))
        rc = lib.wg_taco_decode(h, memory.data_ptr(), tl.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), B, S,
                                int(max_length), int(bool(early_stopping)), int(bool(deterministic)), int(seed),
                                outputs.data_ptr(), stops.data_ptr(), attn.data_ptr() if attn is not None else None,
                                lengths.data_ptr(), ctypes.byref(frames),
                                torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"wg_taco_decode failed ({rc}): {lib.wg_taco_last_error(h).decode()}")
        return outputs, stops, attn, lengths

    def _conv_bn(self, t, conv, bn):
        """Conv1D + inference batch-norm folded into one weight/bias pair."""
        k, b = t(conv + "/kernel"), t(conv + "/bias")                       # [k, in, out]
        scale = t(bn + "/gamma") / torch.sqrt(t(bn + "/moving_variance") + self.hp.bn_epsilon)
        w = (k * scale).permute(2, 1, 0).contiguous()                       # torch layout [out, in, k]
        return w, (b - t(bn + "/moving_mean")) * scale + t(bn + "/beta")

    # ---------------------------------------------------------------- encoder / postnet
    def encode(self, tokens: torch.Tensor):
        """tokens int64 [B, S] (pad_token on the right) -> (memory [B, S, 512] zeroed at pads, mask [B, S])."""
        mask = tokens != self.hp.pad_token
        x = self.emb[tokens] * mask[..., None]
        x = x.transpose(1, 2)
        m = mask[:, None, :]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):    # cuDNN would otherwise run fp32 convs / LSTMs in TF32
            for w, b in self.enc_convs:
                x = F.relu(F.conv1d(x * m, w, b, padding=w.shape[-1] // 2))
            x = (x * m).transpose(1, 2)
            lengths = mask.sum(1).cpu()
            packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True, enforce_sorted=False)
            out, _ = self.bilstm(packed)
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=tokens.shape[1])
        return out * mask[..., None], mask

    def postnet(self, decoder_output: torch.Tensor, mask: torch.Tensor):
        x = (decoder_output * mask[..., None]).transpose(1, 2)
        m = mask[:, None, :]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            for i, (w, b) in enumerate(self.post_convs):
                x = F.conv1d(x * m, w, b, padding=w.shape[-1] // 2)
                if i + 1 < len(self.post_convs):
                    x = torch.tanh(x)
        return (x * m).transpose(1, 2)

    # ---------------------------------------------------------------- decoder
    @staticmethod
    def _lstm(x_h, c, w, b):
        i, f, g, o = (x_h @ w + b).chunk(4, dim=-1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c_new), c_new

    def _step(self, s, memory, pm, neg_mask, deterministic):
        """One decoder frame; `s` holds the recurrent state, updated IN PLACE (graph-capture friendly)."""
        x = s["frame"]
        for w in self.prenet:
            x = F.relu(x @ w)
            if not deterministic:
                x = F.dropout(x, self.hp.prenet_drop_rate, training=True)
        h_a, c_a = self._lstm(torch.cat([x, s["ctx"], s["h_a"]], dim=-1), s["c_a"], self.att_w, self.att_b)
        q = h_a @ self.q_w
        loc = F.conv1d(torch.stack([s["aw"], s["awc"]], dim=1), self.loc_conv, padding=self.loc_conv.shape[-1] // 2)
        loc = loc.transpose(1, 2) @ self.loc_dense
        e = torch.tanh(q[:, None, :] + pm + loc) @ self.v_w + neg_mask
        aw = torch.softmax(e, dim=-1)
        ctx = torch.bmm(aw[:, None, :], memory)[:, 0]
        h_d, c_d = self._lstm(torch.cat([h_a, ctx, s["h_d"]], dim=-1), s["c_d"], self.dec_w, self.dec_b)
        out = torch.cat([h_d, ctx], dim=-1) @ self.out_w + self.out_b
        frame, stop = out[:, :-1], torch.sigmoid(out[:, -1])
        finished = s["finished"] | (stop > 0.5)
        s["lengths"].add_((~finished).to(s["lengths"].dtype))
        s["finished"].copy_(finished)
        s["awc"].add_(aw)
        for k, v in (("frame", frame), ("h_a", h_a), ("c_a", c_a), ("h_d", h_d), ("c_d", c_d), ("ctx", ctx), ("aw", aw)):
            s[k].copy_(v)
        return frame, stop, aw

    def _initial_state(self, B, S):
        hp, z = self.hp, lambda *sh: torch.zeros(*sh, dtype=self.dtype, device=self.device)  # noqa: E731
        return {"frame": z(B, hp.n_mel_channels), "h_a": z(B, hp.attention_rnn_dim), "c_a": z(B, hp.attention_rnn_dim),
                "h_d": z(B, hp.decoder_rnn_dim), "c_d": z(B, hp.decoder_rnn_dim), "ctx": z(B, hp.embedding_dim),
                "aw": z(B, S), "awc": z(B, S),
                "finished": torch.zeros(B, dtype=torch.bool, device=self.device),
                "lengths": torch.zeros(B, dtype=torch.int32, device=self.device)}

    def decode(self, memory, mask, max_length, early_stopping=True, deterministic=False, use_graph=False, graph_chunk=32):
        B, S, _ = memory.shape
        hp = self.hp
        pm = memory @ self.m_w
        neg_mask = torch.zeros(B, S, dtype=self.dtype, device=self.device).masked_fill_(~mask, float("-inf"))
        outputs = torch.zeros(B, max_length, hp.n_mel_channels, dtype=self.dtype, device=self.device)
        stops = torch.zeros(B, max_length, dtype=self.dtype, device=self.device)
        attn = torch.zeros(B, max_length, S, dtype=self.dtype, device=self.device)
        s = self._initial_state(B, S)
        if not use_graph:
            for t in range(max_length):
                if early_stopping and bool(s["finished"].all()):    # K.while_loop cond (:632-634)
                    break
                frame, stop, aw = self._step(s, memory, pm, neg_mask, deterministic)
                outputs[:, t], stops[:, t], attn[:, t] = frame, stop, aw
            return outputs, stops, attn, s["lengths"].clone()
        # CUDA-graph path: `graph_chunk` steps per graph, replayed; state and the chunk outputs are static buffers
        key = (B, S, graph_chunk, bool(deterministic))
        g = self._graphs.get(key)
        if g is None:
            st = self._initial_state(B, S)
            buf = {"memory": torch.zeros_like(memory), "pm": torch.zeros_like(pm), "neg": torch.zeros_like(neg_mask),
                   "frames": torch.zeros(B, graph_chunk, hp.n_mel_channels, dtype=self.dtype, device=self.device),
                   "stops": torch.zeros(B, graph_chunk, dtype=self.dtype, device=self.device),
                   "attn": torch.zeros(B, graph_chunk, S, dtype=self.dtype, device=self.device)}

            def run_chunk():
                for i in range(graph_chunk):
                    frame, stop, aw = self._step(st, buf["memory"], buf["pm"], buf["neg"], deterministic)
                    buf["frames"][:, i], buf["stops"][:, i], buf["attn"][:, i] = frame, stop, aw

            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                run_chunk()                              # warm-up outside capture (cuBLAS workspaces, lazy init)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run_chunk()
            while len(self._graphs) >= self.max_graphs:          # LRU: each entry pins static state / memory buffers
                self._graphs.pop(next(iter(self._graphs)))
            g = (graph, st, buf)
        else:
            self._graphs.pop(key)
        self._graphs[key] = g                                     # most recently used last
        graph, st, buf = g
        for k, v in s.items():
            st[k].copy_(v)
        buf["memory"].copy_(memory), buf["pm"].copy_(pm), buf["neg"].copy_(neg_mask)
        t = 0
        while t < max_length:
            if early_stopping and bool(st["finished"].all()):
                break
            graph.replay()
            n = min(graph_chunk, max_length - t)
            outputs[:, t:t + n], stops[:, t:t + n], attn[:, t:t + n] = buf["frames"][:, :n], buf["stops"][:, :n], buf["attn"][:, :n]
            t += n
        lengths = st["lengths"].clone()
        if t < max_length or max_length % graph_chunk:
            # a chunk may run past max_length / past the step at which everything had finished:
            # lengths only grow while an utterance is unfinished, so clamping restores the loop's value
            lengths.clamp_(max=max_length)
        return outputs, stops, attn, lengths

    @torch.no_grad()
    def infer(self, inputs, *, max_length=None, early_stopping=True, deterministic=False, use_graph=False,
              graph_chunk=32, decoder="torch", seed=0, return_attention=True, **_):
        """`inputs`: int tokens [B, S]; `max_length`: int frames, or float = frames per token of the longest
        text (tacotron2_arch.py:886-892). Returns Tacotron2InferenceOutput (torch tensors on the device)."""
        tokens = torch.as_tensor(np.asarray(inputs) if not isinstance(inputs, torch.Tensor) else inputs).to(self.device).long()
        if tokens.dim() == 1:
            tokens = tokens[None]
        memory, mask = self.encode(tokens)
        if max_length is None:
            raise ValueError("max_length is required (the reference's max_decoder_steps default is None too)")
        if isinstance(max_length, float):
            max_length = int(float(mask.sum(1).max()) * max_length)
        if decoder == "b200":
            dec, stops, attn, lengths = self.decode_b200(memory, mask, int(max_length), early_stopping, deterministic,
                                                         seed, graph_chunk, return_attention)
        elif decoder == "torch":
            dec, stops, attn, lengths = self.decode(memory, mask, int(max_length), early_stopping, deterministic,
                                                    use_graph, graph_chunk)
        else:
            raise ValueError(f"decoder must be 'torch' or 'b200', got {decoder!r}")
        dmask = torch.arange(dec.shape[1], device=self.device)[None] <= lengths[:, None]
        mel = dec + self.postnet(dec, dmask)
        return Tacotron2InferenceOutput(decoder_output=dec, mel=mel, stop_tokens=stops, attention_weights=attn,
                                        lengths=lengths)
